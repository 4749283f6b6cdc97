"""`from concrete.ml.sklearn.base import QuantizedModule` (/root/reference/fhe_similarity.py:6): imported by the
reference, never used."""


class QuantizedModule:  # pragma: no cover - a name to import, nothing more
    pass
