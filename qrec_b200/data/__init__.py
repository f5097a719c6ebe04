"""qrec_b200.data: engine-backed mirror of the reference package of the same name."""
