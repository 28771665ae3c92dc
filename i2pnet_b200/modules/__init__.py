"""Mirror of the parts of the reference's src/modules the large-range path uses."""
