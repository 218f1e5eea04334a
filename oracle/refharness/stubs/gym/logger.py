def set_level(level):
    pass
