"""Reference import path `models.*` served by the B200 backend (see pde_surrogate_b200)."""
