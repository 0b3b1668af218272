"""B200-native stable-fluids step — host-side package.

Holds only what the hot path needs: ``csrc/`` (sm_100a kernels + the C ABI of
``include/fluid_b200.h``), the ctypes binding, and a mirror of the reference's
operator surface (``advect``, ``calculate_divergence``, ``subtract_gradient``,
``poisson_solve``).  Importing the operators requires the built CUDA library;
there is no CPU fallback.
"""
from .build import LIB_PATH, build_cuda, build_harness  # noqa: F401

__all__ = ["LIB_PATH", "build_cuda", "build_harness", "Context", "Tile", "FluidError",
           "advect", "calculate_divergence", "subtract_gradient", "poisson_solve", "DRAG_DTYPE"]


def __getattr__(name):
    # lazy: `import esp32_fluid_simulation_b200` must work before the library is built
    if name in ("Context", "Sim", "advect", "calculate_divergence", "subtract_gradient", "poisson_solve",
                "DRAG_DTYPE", "default_context"):
        from . import ops
        return getattr(ops, name)
    if name in ("Tile", "FluidError"):
        from . import _lib
        return getattr(_lib, name)
    raise AttributeError(name)
