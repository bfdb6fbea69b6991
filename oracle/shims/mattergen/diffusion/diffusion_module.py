class DiffusionModule:
    pass
