"""Drop-in alias of the reference's ``safe_pilco_extension`` package (implemented in ``pilco_b200.safe``)."""
