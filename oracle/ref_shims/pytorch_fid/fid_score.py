def calculate_frechet_distance(*a, **k):
    raise NotImplementedError
