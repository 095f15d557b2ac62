def makedirs(*a, **k):
    return None
