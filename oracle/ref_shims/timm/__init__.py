"""Import shim (test infrastructure) for `timm` -- only two helpers are used by the reference."""
