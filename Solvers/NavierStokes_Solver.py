from sem_b200.NavierStokes_Solver import NavierStokesSolver  # noqa: F401
