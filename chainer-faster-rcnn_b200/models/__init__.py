"""Drop-in mirror of the reference's `models` package (same module paths, class names, call
signatures) whose arithmetic runs entirely in libfrcnn_b200.so on a B200.  See INTEGRATION.md."""
