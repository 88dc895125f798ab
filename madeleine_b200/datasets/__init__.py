"""Input side of the hot path (drop-in for madeleine/datasets): the reference's CPU datasets / collate and the
HBM-resident feature store with on-device resampling."""
