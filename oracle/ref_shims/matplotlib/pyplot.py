"""Shim: only imported, never called, by the reference code paths the goldens exercise."""
