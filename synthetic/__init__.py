"""Synthetic inputs: what Firedrake hands to the solver in the reference (assembled spatial
matrices on structured meshes, tested data vectors).  Neither product nor oracle: the
benchmarks, the tests and the oracle all take their inputs from here, so that every
implementation sees the same matrices.  Nothing in here solves anything."""
