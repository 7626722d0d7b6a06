"""Drop-in mirror of the reference's `module` package for the hot path (module/srvp.py, conv.py, mlp.py, utils.py)."""
