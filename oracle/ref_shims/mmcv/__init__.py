"""Import shim (test infrastructure) for `mmcv` -- layer builders only (reference: lib/models/hrformer.py:7-22)."""
