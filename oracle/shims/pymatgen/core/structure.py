class Structure:     # type-hint only in memory/ltm.py of the reference; tests pass duck-typed stand-ins
    pass
