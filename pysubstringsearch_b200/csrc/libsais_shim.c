/* libsais_shim.c — optional shim: exports the reference's own foreign symbol `libsais`
 * (src/libsais/libsais.h:56-65, declared in src/lib.rs:14-22) and forwards it to pss_libsais, so
 * that the reference's Rust host needs NO source change at all: build.rs stops compiling
 * src/libsais/libsais.c and links  -lsais_pss_shim -lpss_b200  instead.
 * Same arguments, same ownership, same return codes (0 / -1 / -2, plus -3 when CUDA fails).
 * Kept out of libpss_b200 itself so that a process which also carries the real libsais never
 * sees two definitions of the symbol. */
#include "../../include/pss.h"

#ifdef __cplusplus
extern "C"
#endif
int32_t libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq) {
    return pss_libsais(T, SA, n, fs, freq);
}
