"""Mirror of the reference's ``lib`` package surface for the CRNN+CTC hot path."""
