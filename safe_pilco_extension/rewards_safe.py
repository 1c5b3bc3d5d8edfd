"""``from safe_pilco_extension.rewards_safe import RiskOfCollision, SingleConstraint, ObjectiveFunction``
(reference: safe_pilco_extension/rewards_safe.py) -- implemented on the device in ``pilco_b200.safe``."""
from pilco_b200.safe import RiskOfCollision, SingleConstraint, ObjectiveFunction   # noqa: F401
