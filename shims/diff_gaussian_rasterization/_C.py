"""Drop-in for upstream's pybind11 module ``diff_gaussian_rasterization._C`` ("C++/FFI surface"): the
three functions with upstream's positional signatures and return tuples, for code that calls the extension below the
``GaussianRasterizer`` module.  The byte buffers they exchange are manus_b200's opaque forward state."""
from manus_b200.rasterizer import c_mark_visible as mark_visible  # noqa: F401
from manus_b200.rasterizer import c_rasterize_gaussians as rasterize_gaussians  # noqa: F401
from manus_b200.rasterizer import c_rasterize_gaussians_backward as rasterize_gaussians_backward  # noqa: F401
