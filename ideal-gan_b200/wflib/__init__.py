from wflib.IDEAL_model import *  # noqa: F401,F403  (same import surface as the reference's wflib/__init__.py:1)
