class UPValueError(Exception):
    pass
