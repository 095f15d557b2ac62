from typing import Any
IndexType = Any


class Dataset:
    pass
