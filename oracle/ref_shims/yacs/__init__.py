"""Import shim (test infrastructure): stands in for `yacs`, which is absent from this image.
Only what the reference's lib/config/default.py uses."""
