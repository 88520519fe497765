"""Import path of the reference's vivsim/lbm/forcing/ package; the functions live in vivsim_b200.lbm."""
