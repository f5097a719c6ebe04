"""qrec_b200: B200-native engine behind QRec's BPR / LightGCN / NeuMF hot path."""
