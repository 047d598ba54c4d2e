from sem_b200.ConvectionDiffusion_Solver import ConvectionDiffusionSolver  # noqa: F401
