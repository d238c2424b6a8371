"""Reference import path `utils.*`: SobelFilter on the B200 backend plus the host-side harness
helpers the training script imports (data loading, LR schedule, stats/plots)."""
