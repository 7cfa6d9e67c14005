"""hostgen -- TEST / BENCH HARNESS, not part of the product.

The product (bolt.jl_b200/, libbolt_cuda.so) takes the 1-D background and the RECFAST ionization history as spline tables computed
on the HOST by the reference's own Julia code (north star: "the cheap 1-D background and RECFAST/ionization history stay on the
host").  Julia does not exist in the build image or on the GPU box, so the tests and the benchmark need some other producer of
those tables: this package restates the reference's out-of-scope host components for that purpose only --
src/background.jl:5-128 (background.py), src/ionization/recfast.jl + ionization.jl (recfast.py), Interpolations.jl's cubic B-spline
prefilter (bspline.py), the unit constants of Unitful/PhysicalConstants (constants.py).  It is pinned by the reference's Fortran
RECFAST fixture (tests/test_host_inputs.py) and earns no coverage credit.  `batch.py` is the batched generator.
"""
from bolt_b200.params import CosmoParams                        # noqa: F401  (the product's parameter container)
from .background import Background                              # noqa: F401
from .recfast import RECFAST, IonizationHistory                 # noqa: F401
from .partials import host_cosmo_with_partials                  # noqa: F401
