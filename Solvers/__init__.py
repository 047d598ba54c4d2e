"""Import-path shim: the reference's callers do ``from Solvers.NavierStokes_Solver import NavierStokesSolver``
(OpenMDAO/Boussinesq_SequentialCoupler.py:3-4, Examples/*.py).  With this repository on ``sys.path`` ahead of the
reference, those imports resolve to the GPU drop-in classes of ``sem_b200``."""
