"""qrec_b200.util: engine-backed mirror of the reference package of the same name."""
