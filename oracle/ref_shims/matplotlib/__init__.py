"""Import shim (TEST INFRASTRUCTURE): lib/dataset/JointsDataset.py imports matplotlib.pyplot for a debug viewer only."""
