"""TEST INFRASTRUCTURE: a stand-in for the un-vendored `concrete` package, so that the reference's own scripts
(/root/reference/*.py, unmodified) import this repo's estimators where they import Concrete-ML's.  See
tests/ref_shim/concrete/ml/sklearn/__init__.py."""
