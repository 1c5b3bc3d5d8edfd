"""``from safe_pilco_extension.safe_pilco import SafePILCO`` (reference: safe_pilco_extension/safe_pilco.py)
-- implemented on the device in ``pilco_b200.safe``."""
from pilco_b200.safe import SafePILCO   # noqa: F401
