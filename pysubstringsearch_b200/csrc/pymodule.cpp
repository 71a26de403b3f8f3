// pymodule.cpp — CPython extension `pysubstringsearch_b200.pysubstringsearch`: the native
// classes the Python façade forwards to by keyword, with the parameter names, defaults and
// exception types of the reference's pyo3 module (src/lib.rs:42-136 Writer, :155-288 Reader,
// :290-299 module; façade calls at pysubstringsearch/__init__.py:12-15, 21-23, 29-31, 49-51,
// 57-59).  Everything below the binding goes through the C ABI of include/pss.h.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cstring>
#include <string>
#include <vector>

#include "pss.h"

namespace {

PyObject *raise_status(int rc) {
    const char *msg = pss_last_error();
    switch (rc) {
        case PSS_ERR_NOTFOUND: PyErr_SetString(PyExc_FileNotFoundError, msg); break;
        case PSS_ERR_IO:       PyErr_SetString(PyExc_OSError, msg); break;
        case PSS_ERR_FORMAT:   PyErr_SetString(PyExc_OSError, msg); break;
        case PSS_ERR_TOOBIG:   PyErr_SetString(PyExc_ValueError, "entry is too big"); break;
        case PSS_ERR_ARG:      PyErr_SetString(PyExc_ValueError, msg); break;
        case PSS_ERR_NOMEM:    PyErr_SetString(PyExc_MemoryError, msg); break;
        default:               PyErr_SetString(PyExc_RuntimeError, msg); break;
    }
    return nullptr;
}

// ---- Writer ---------------------------------------------------------------------------
// pyo3 hands out `&mut self` for every method of the reference's classes (lib.rs:67,88,105,
// 126,201): a second thread entering while a call is in flight gets RuntimeError("Already
// borrowed").  The GIL is released around the native calls here, so the same exclusivity
// is enforced explicitly (the C handles are not thread-safe).
struct BorrowGuard {
    bool *flag;
    bool  ok;
    explicit BorrowGuard(bool *f) : flag(f), ok(!*f) {
        if (ok) *flag = true;
        else PyErr_SetString(PyExc_RuntimeError, "Already borrowed");
    }
    ~BorrowGuard() { if (ok) *flag = false; }
};

struct WriterObject {
    PyObject_HEAD
    pss_writer *w;
    bool busy;
};

int Writer_init(WriterObject *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"index_file_path", "max_chunk_len", nullptr};
    PyObject *path_obj = nullptr, *max_obj = Py_None;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "U|O", const_cast<char **>(kwlist), &path_obj, &max_obj)) return -1;
    long long max_chunk_len = -1;
    if (max_obj != Py_None) {
        max_chunk_len = PyLong_AsLongLong(max_obj);
        if (max_chunk_len == -1 && PyErr_Occurred()) return -1;
        if (max_chunk_len < 0) {
            PyErr_SetString(PyExc_OverflowError, "can't convert negative int to unsigned");
            return -1;
        }
    }
    const char *path = PyUnicode_AsUTF8(path_obj);
    if (!path) return -1;
    if (self->w) { pss_writer_close(self->w); self->w = nullptr; }
    int rc = pss_writer_open(path, max_chunk_len, &self->w);
    if (rc != PSS_OK) { raise_status(rc); return -1; }
    return 0;
}

void Writer_dealloc(WriterObject *self) {
    if (self->w) {
        // Drop (lib.rs:138-144): flush what is buffered.  A failure cannot be raised from a
        // destructor; report it without aborting the interpreter.
        PyObject *et, *ev, *tb;
        PyErr_Fetch(&et, &ev, &tb);
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = pss_writer_close(self->w);
        Py_END_ALLOW_THREADS
        self->w = nullptr;
        if (rc != PSS_OK) {
            raise_status(rc);
            PyErr_WriteUnraisable(nullptr);  // not `self`: it is already being torn down
        }
        PyErr_Restore(et, ev, tb);
    }
    Py_TYPE(self)->tp_free(reinterpret_cast<PyObject *>(self));
}

// add_entry is called once per entry (11 M times for a 500 MB corpus): vectorcall convention, and
// the two well-formed call shapes — add_entry(s) and add_entry(text=s), the façade's — are
// recognised without building a tuple and a dict (446 -> ~250 ns per call here).  Anything else
// takes the generic parser, which produces the usual TypeError texts.
PyObject *Writer_add_entry(WriterObject *self, PyObject *const *args, Py_ssize_t nargs, PyObject *kwnames) {
    PyObject *text = nullptr;
    const Py_ssize_t nkw = kwnames ? PyTuple_GET_SIZE(kwnames) : 0;
    if (nargs == 1 && nkw == 0) {
        text = args[0];
    } else if (nargs == 0 && nkw == 1 && PyUnicode_CompareWithASCIIString(PyTuple_GET_ITEM(kwnames, 0), "text") == 0) {
        text = args[0];
    }
    if (!text || !PyUnicode_Check(text)) {
        static const char *kwlist[] = {"text", nullptr};
        PyObject *tuple = PyTuple_New(nargs), *dict = nkw ? PyDict_New() : nullptr;
        if (!tuple || (nkw && !dict)) { Py_XDECREF(tuple); Py_XDECREF(dict); return nullptr; }
        for (Py_ssize_t i = 0; i < nargs; ++i) { Py_INCREF(args[i]); PyTuple_SET_ITEM(tuple, i, args[i]); }
        bool ok = true;
        for (Py_ssize_t i = 0; i < nkw && ok; ++i) ok = PyDict_SetItem(dict, PyTuple_GET_ITEM(kwnames, i), args[nargs + i]) == 0;
        text = nullptr;
        ok = ok && PyArg_ParseTupleAndKeywords(tuple, dict, "U", const_cast<char **>(kwlist), &text);
        Py_DECREF(tuple);
        Py_XDECREF(dict);
        if (!ok) return nullptr;      // `text` is borrowed from the caller's argument array, which outlives the call
    }
    Py_ssize_t len = 0;
    const char *p = PyUnicode_AsUTF8AndSize(text, &len);
    if (!p) return nullptr;
    BorrowGuard guard(&self->busy);
    if (!guard.ok) return nullptr;
    int rc;
    if (pss_writer_would_flush(self->w, (size_t)len)) {
        // this entry hands the buffered chunk to the GPU builder and may wait for a free
        // pipeline slot: other Python threads run meanwhile.  `p` stays valid: `text` is
        // referenced by the caller's argument tuple, and BorrowGuard keeps the handle ours.
        Py_BEGIN_ALLOW_THREADS
        rc = pss_writer_add_entry(self->w, reinterpret_cast<const uint8_t *>(p), (size_t)len);
        Py_END_ALLOW_THREADS
    } else {
        rc = pss_writer_add_entry(self->w, reinterpret_cast<const uint8_t *>(p), (size_t)len);
    }
    if (rc != PSS_OK) return raise_status(rc);
    Py_RETURN_NONE;
}

PyObject *Writer_add_entries_from_file_lines(WriterObject *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"input_file_path", nullptr};
    PyObject *path_obj = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "U", const_cast<char **>(kwlist), &path_obj)) return nullptr;
    const char *path = PyUnicode_AsUTF8(path_obj);
    if (!path) return nullptr;
    std::string path_copy(path);
    BorrowGuard guard(&self->busy);
    if (!guard.ok) return nullptr;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = pss_writer_add_entries_from_file_lines(self->w, path_copy.c_str());
    Py_END_ALLOW_THREADS
    if (rc != PSS_OK) return raise_status(rc);
    Py_RETURN_NONE;
}

PyObject *Writer_dump_data(WriterObject *self, PyObject *) {
    BorrowGuard guard(&self->busy);
    if (!guard.ok) return nullptr;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = pss_writer_dump_data(self->w);
    Py_END_ALLOW_THREADS
    if (rc != PSS_OK) return raise_status(rc);
    Py_RETURN_NONE;
}

PyObject *Writer_finalize(WriterObject *self, PyObject *) {
    BorrowGuard guard(&self->busy);
    if (!guard.ok) return nullptr;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = pss_writer_finalize(self->w);
    Py_END_ALLOW_THREADS
    if (rc != PSS_OK) return raise_status(rc);
    Py_RETURN_NONE;
}

PyMethodDef Writer_methods[] = {
    {"add_entries_from_file_lines", reinterpret_cast<PyCFunction>(Writer_add_entries_from_file_lines),
     METH_VARARGS | METH_KEYWORDS, "add_entries_from_file_lines(input_file_path)"},
    {"add_entry", reinterpret_cast<PyCFunction>(reinterpret_cast<void (*)(void)>(Writer_add_entry)),
     METH_FASTCALL | METH_KEYWORDS, "add_entry(text)"},
    {"dump_data", reinterpret_cast<PyCFunction>(Writer_dump_data), METH_NOARGS, "dump_data()"},
    {"finalize", reinterpret_cast<PyCFunction>(Writer_finalize), METH_NOARGS, "finalize()"},
    {nullptr, nullptr, 0, nullptr}};

PyTypeObject WriterType = {PyVarObject_HEAD_INIT(nullptr, 0)};

// ---- Reader ---------------------------------------------------------------------------
struct ReaderObject {
    PyObject_HEAD
    pss_reader *r;
    bool busy;
};

int Reader_init(ReaderObject *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"index_file_path", nullptr};
    PyObject *path_obj = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "U", const_cast<char **>(kwlist), &path_obj)) return -1;
    const char *path = PyUnicode_AsUTF8(path_obj);
    if (!path) return -1;
    std::string path_copy(path);
    if (self->r) { pss_reader_close(self->r); self->r = nullptr; }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = pss_reader_open(path_copy.c_str(), &self->r);
    Py_END_ALLOW_THREADS
    if (rc != PSS_OK) { raise_status(rc); return -1; }
    return 0;
}

void Reader_dealloc(ReaderObject *self) {
    if (self->r) {
        pss_reader *r = self->r;
        self->r = nullptr;
        Py_BEGIN_ALLOW_THREADS      // frees gigabytes of GPU and host memory
        pss_reader_close(r);
        Py_END_ALLOW_THREADS
    }
    Py_TYPE(self)->tp_free(reinterpret_cast<PyObject *>(self));
}

// str from entry bytes.  Most corpora are ASCII: one pass of 8-byte words checks that, and a
// 1-byte-kind str is filled with memcpy — about half the cost of the general UTF-8 decoder.
inline PyObject *make_str(const uint8_t *p, size_t n) {
    uint64_t acc = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        acc |= w;
    }
    for (; i < n; ++i) acc |= p[i];
    if ((acc & 0x8080808080808080ull) == 0) {
        PyObject *o = PyUnicode_New((Py_ssize_t)n, 127);
        if (o && n) std::memcpy(PyUnicode_1BYTE_DATA(o), p, n);
        return o;
    }
    // The reference hands out from_utf8_unchecked slices (lib.rs:275); entries were added as
    // valid UTF-8, so strict decoding succeeds; foreign bytes decode with surrogates.
    return PyUnicode_DecodeUTF8(reinterpret_cast<const char *>(p), (Py_ssize_t)n, "surrogateescape");
}

// One batched native search; returns the concatenated list of entries (query order).
PyObject *run_batch(ReaderObject *self, const std::vector<uint8_t> &blob, const std::vector<int64_t> &offsets) {
    BorrowGuard guard(&self->busy);
    if (!guard.ok) return nullptr;
    pss_result *res = nullptr;
    int rc;
    const int32_t nq = (int32_t)offsets.size() - 1;
    Py_BEGIN_ALLOW_THREADS
    rc = pss_reader_search_batch(self->r, blob.data(), offsets.data(), nq, &res);
    Py_END_ALLOW_THREADS
    if (rc != PSS_OK) return raise_status(rc);
    PyObject *list = PyList_New((Py_ssize_t)res->n_entries);
    if (!list) { pss_result_free(res); return nullptr; }
    const uint8_t *text = nullptr;
    int64_t text_len = 0;
    int32_t cur_chunk = -1;
    // Entries sit at random offsets of a multi-hundred-MB text: without prefetching, every
    // string costs a DRAM miss on the host.  Touch the lines of the entry 16 ahead.
    constexpr int64_t AHEAD = 16;
    for (int64_t i = 0; i < res->n_entries; ++i) {
        if (i + AHEAD < res->n_entries && res->chunk_id[i + AHEAD] == cur_chunk && text) {
            const uint8_t *pf = text + res->line_start[i + AHEAD];
            __builtin_prefetch(pf);
            __builtin_prefetch(pf + 64);
        }
        if (res->chunk_id[i] != cur_chunk) {
            cur_chunk = res->chunk_id[i];
            pss_reader_chunk_text(self->r, cur_chunk, &text, &text_len);
        }
        const uint32_t s = res->line_start[i], e = res->line_end[i];
        PyObject *str = make_str(text + s, (size_t)(e - s));
        if (!str) { Py_DECREF(list); pss_result_free(res); return nullptr; }
        PyList_SET_ITEM(list, (Py_ssize_t)i, str);
    }
    pss_result_free(res);
    return list;
}

PyObject *Reader_search(ReaderObject *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"substring", nullptr};
    PyObject *sub = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "U", const_cast<char **>(kwlist), &sub)) return nullptr;
    Py_ssize_t len = 0;
    const char *p = PyUnicode_AsUTF8AndSize(sub, &len);
    if (!p) return nullptr;
    std::vector<uint8_t> blob(reinterpret_cast<const uint8_t *>(p), reinterpret_cast<const uint8_t *>(p) + len);
    std::vector<int64_t> offsets = {0, (int64_t)len};
    return run_batch(self, blob, offsets);
}

PyObject *Reader_search_multiple(ReaderObject *self, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"substrings", nullptr};
    PyObject *subs = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "O", const_cast<char **>(kwlist), &subs)) return nullptr;
    PyObject *seq = PySequence_Fast(subs, "substrings must be a sequence of str");
    if (!seq) return nullptr;
    const Py_ssize_t nq = PySequence_Fast_GET_SIZE(seq);
    std::vector<uint8_t> blob;
    std::vector<int64_t> offsets;
    offsets.reserve((size_t)nq + 1);
    offsets.push_back(0);
    for (Py_ssize_t i = 0; i < nq; ++i) {
        PyObject *item = PySequence_Fast_GET_ITEM(seq, i);
        if (!PyUnicode_Check(item)) {
            Py_DECREF(seq);
            PyErr_Format(PyExc_TypeError, "argument 'substring': '%.100s' object cannot be converted to 'PyString'",
                         Py_TYPE(item)->tp_name);
            return nullptr;
        }
        Py_ssize_t len = 0;
        const char *p = PyUnicode_AsUTF8AndSize(item, &len);
        if (!p) { Py_DECREF(seq); return nullptr; }
        blob.insert(blob.end(), reinterpret_cast<const uint8_t *>(p), reinterpret_cast<const uint8_t *>(p) + len);
        offsets.push_back((int64_t)blob.size());
    }
    Py_DECREF(seq);
    return run_batch(self, blob, offsets);
}

PyMethodDef Reader_methods[] = {
    {"search", reinterpret_cast<PyCFunction>(Reader_search), METH_VARARGS | METH_KEYWORDS, "search(substring) -> list[str]"},
    {"search_multiple", reinterpret_cast<PyCFunction>(Reader_search_multiple), METH_VARARGS | METH_KEYWORDS,
     "search_multiple(substrings) -> list[str]  (one batched GPU call)"},
    {nullptr, nullptr, 0, nullptr}};

PyTypeObject ReaderType = {PyVarObject_HEAD_INIT(nullptr, 0)};

PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "pysubstringsearch",
                         "B200-native Writer/Reader (suffix-array substring search)", -1, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit_pysubstringsearch(void) {
    WriterType.tp_name      = "pysubstringsearch.Writer";
    WriterType.tp_basicsize = sizeof(WriterObject);
    WriterType.tp_flags     = Py_TPFLAGS_DEFAULT;
    WriterType.tp_new       = PyType_GenericNew;
    WriterType.tp_init      = reinterpret_cast<initproc>(Writer_init);
    WriterType.tp_dealloc   = reinterpret_cast<destructor>(Writer_dealloc);
    WriterType.tp_methods   = Writer_methods;
    ReaderType.tp_name      = "pysubstringsearch.Reader";
    ReaderType.tp_basicsize = sizeof(ReaderObject);
    ReaderType.tp_flags     = Py_TPFLAGS_DEFAULT;
    ReaderType.tp_new       = PyType_GenericNew;
    ReaderType.tp_init      = reinterpret_cast<initproc>(Reader_init);
    ReaderType.tp_dealloc   = reinterpret_cast<destructor>(Reader_dealloc);
    ReaderType.tp_methods   = Reader_methods;
    if (PyType_Ready(&WriterType) < 0 || PyType_Ready(&ReaderType) < 0) return nullptr;
    PyObject *m = PyModule_Create(&moduledef);
    if (!m) return nullptr;
    Py_INCREF(&WriterType);
    Py_INCREF(&ReaderType);
    if (PyModule_AddObject(m, "Writer", reinterpret_cast<PyObject *>(&WriterType)) < 0 ||
        PyModule_AddObject(m, "Reader", reinterpret_cast<PyObject *>(&ReaderType)) < 0) {
        Py_DECREF(m);
        return nullptr;
    }
    PyModule_AddStringConstant(m, "__backend__", pss_version());
    return m;
}
