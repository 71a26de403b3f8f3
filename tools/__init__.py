"""Developer tools: synthetic corpora, ncu drivers, launch-list summaries."""
