"""Stand-in for `pytorch3d`: only `pytorch3d.transforms.quaternion_to_matrix` (scene/gaussian_model.py:23,163)."""
