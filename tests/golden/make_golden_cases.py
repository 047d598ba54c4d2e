"""Case tables shared by make_golden.py (generator) and the tests (consumers)."""

MESHES = [  # (tag, P, N_ex, N_ey, L_x, L_y)
    ("p4_8x8", 4, 8, 8, 1.0, 1.0),
    ("p4_16x16", 4, 16, 16, 1.0, 1.0),
    ("p3_3x5", 3, 3, 5, 2.0, 1.5),
    ("p8_4x3", 8, 4, 3, 1.0, 0.7),
    ("p1_6x4", 1, 6, 4, 1.0, 1.0),
    ("p2_1x1", 2, 1, 1, 1.0, 1.0),
    ("p5_7x2", 5, 7, 2, 3.0, 1.0),
]

CD_CASES = [  # tag, kwargs
    ("c1", dict(L_x=1, L_y=1, Pe=40, P=4, N_ex=16, N_ey=16, T_E=-0.5, T_W=0.5)),
    ("allbc", dict(L_x=2.0, L_y=1.5, Pe=7.5, P=3, N_ex=3, N_ey=5, T_W=0.3, T_E=-0.2, T_S=1.0, T_N=0.0)),
    ("sn", dict(L_x=1.0, L_y=0.7, Pe=12.0, P=8, N_ex=4, N_ey=3, T_S=0.25, T_N=-1.0)),
    ("p5", dict(L_x=3.0, L_y=1.0, Pe=3.0, P=5, N_ex=7, N_ey=2, T_W=1.0)),
]

NS_CASES = [
    ("small", dict(L_x=1.0, L_y=1.0, Re=50.0, Gr=200.0, P=3, N_ex=4, N_ey=4, u_N=1.0), True),
    # pin node int(N/2) on the boundary (N even): the reference's own Newton solve fails there, applies only
    ("pinbnd", dict(L_x=2.0, L_y=1.0, Re=20.0, Gr=0.0, P=3, N_ex=3, N_ey=5, v_W=0.3, v_E=-0.2, u_S=0.1, u_N=1.0), False),
    ("rect", dict(L_x=2.0, L_y=1.0, Re=20.0, Gr=30.0, P=2, N_ex=3, N_ey=5, v_W=0.3, v_E=-0.2, u_S=0.1, u_N=1.0), True),
    ("c3", dict(L_x=1.0, L_y=1.0, Re=1e3, Gr=1e3 / 0.71, P=4, N_ex=8, N_ey=8), True),
    ("c2", dict(L_x=1, L_y=1, Re=400, Gr=0, P=4, N_ex=16, N_ey=16, u_N=1), True),
    ("p8", dict(L_x=1.0, L_y=0.7, Re=10.0, Gr=5.0, P=8, N_ex=4, N_ey=3, u_N=1.0), False),
]
