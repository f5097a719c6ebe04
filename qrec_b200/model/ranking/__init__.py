"""qrec_b200.model.ranking: engine-backed mirror of the reference package of the same name."""
