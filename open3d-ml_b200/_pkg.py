"""Package body of open3d_ml_b200 (see open3d_ml_b200/__init__.py)."""
__version__ = "0.1.0"
