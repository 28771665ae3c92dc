"""Mirror of the reference's `pointnet2` package surface (pointnet2/pointnet2_utils.py)."""
