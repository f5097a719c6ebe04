"""qrec_b200.base: engine-backed mirror of the reference package of the same name."""
