"""PyTorch-facing half of the drop-in package: `models` (BaseModel + the RAT_m0..m3 classes on the rat_native engine),
`data_generator` (host DataLoader mirror and the HBM-resident generator) and `torch_utils` (seeding / device helpers)."""
