def compute_prdc(*a, **k):
    raise NotImplementedError
