"""Mirror of the reference's src/projectPN package (utils.py, PPBackbone_center.py)."""
