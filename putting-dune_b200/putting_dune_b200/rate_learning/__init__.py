"""Forward path of the learned rate model (reference:
putting_dune/rate_learning).  Training, distillation and dataset generation
are offline fitting and out of scope."""
