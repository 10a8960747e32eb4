"""Serial stand-in for mpi4py, used ONLY by tools/ to run the reference pylbm in
the build container (mpi4py is not installed there). Carries no arithmetic."""
