"""qrec_b200.model: engine-backed mirror of the reference package of the same name."""
