"""Minimal stand-in for the third-party package ``openmdao`` (pinned ~=3.9.2 by the reference's requirements.txt:4; not installed in
this image, no network).  TEST INFRASTRUCTURE: it exists so that the reference's UNMODIFIED adapters
(``OpenMDAO/*_Component.py``) and coupler script (``OpenMDAO/Boussinesq_SequentialCoupler.py``) can be imported and driven --
over the reference's solvers to record a call trace (tests/golden/make_component_trace.py), and over the drop-in solvers in the
tests.  It restates only the API surface those files touch (see api.py); it is not a re-implementation of OpenMDAO and makes no
claim of iteration-level parity with it."""
