"""The reference's own test files (unmodified).  ``tests/conftest.py`` marks them ``gpu`` (the drop-in classes have
no CPU path) and puts ``tests/shims`` (oct2py / gpflow / tensorflow stand-ins) on the import path."""
