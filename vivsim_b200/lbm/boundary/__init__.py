"""Import path of the reference's vivsim/lbm/boundary/ package; the functions live in vivsim_b200.lbm."""
