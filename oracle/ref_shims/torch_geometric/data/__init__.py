class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class InMemoryDataset:
    def __init__(self, *a, **k):
        pass


class Batch(Data):
    pass
