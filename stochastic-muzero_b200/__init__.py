# Sources of the `stochastic_muzero_b200` package (see ../stochastic_muzero_b200/__init__.py, which
# makes this directory importable under a valid module name).
