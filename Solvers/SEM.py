from sem_b200.SEM import *  # noqa: F401,F403
from sem_b200.SEM import x2xi, xi2x  # noqa: F401
