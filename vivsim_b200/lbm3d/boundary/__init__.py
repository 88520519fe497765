"""Import path of the reference's vivsim/lbm3d/boundary/ package; the functions live in vivsim_b200.lbm3d."""
