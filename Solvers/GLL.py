from sem_b200.GLL import *  # noqa: F401,F403
