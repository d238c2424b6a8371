def switch_backend(*args, **kwargs):
    return None
