"""Linear operators with non-ZERO boundaries / piecewise-constant coefficients, written as a PhiML user writes them.  Shared by
tests/golden/make_golden.py (runs them on the reference's own backends) and tests/test_gpu_phiml.py (runs them through the plug-in).
PhiML is imported lazily: the module itself is importable without the reference."""
import numpy as np

DIMS = 'xyz'
VELOCITY = (1.0, 0.5, 0.25)


def _mixed_ext(rank):
    from phiml.math import extrapolation
    sides = {DIMS[0]: extrapolation.ZERO}
    if rank > 1:
        sides[DIMS[1]] = extrapolation.ZERO_GRADIENT
    if rank > 2:
        sides[DIMS[2]] = extrapolation.PERIODIC
    return extrapolation.combine_sides(**sides)


def _coef_for(x):
    """piecewise-constant coefficient field: 1 in the lower half along the first dim, 3 in the upper half, 0.5 in one corner block"""
    from phiml import math
    sizes = x.shape.spatial.sizes
    k = np.ones(sizes)
    k[sizes[0] // 2:] = 3.0
    k[tuple(slice(0, max(1, s // 3)) for s in sizes)] = 0.5
    return math.tensor(k, x.shape.spatial)


def neumann(x):
    from phiml import math
    return 0.05 * x - math.laplace(x, padding=math.extrapolation.ZERO_GRADIENT)


def periodic(x):
    from phiml import math
    return 0.05 * x - math.laplace(x, padding=math.extrapolation.PERIODIC)


def mixed(x):
    from phiml import math
    return -math.laplace(x, padding=_mixed_ext(x.shape.spatial_rank))


def varcoef(x):
    from phiml import math
    return x - _coef_for(x) * math.laplace(x, padding=math.extrapolation.ZERO)


def periodic_advdiff(x):
    from phiml import math
    from phiml.math import channel
    dims = ','.join(DIMS[:x.shape.spatial_rank])
    g = math.spatial_gradient(x, padding=math.extrapolation.PERIODIC, difference='central', stack_dim=channel(gradient=dims))
    u = math.wrap(list(VELOCITY[:x.shape.spatial_rank]), channel(gradient=dims))
    return x + 0.3 * math.sum(g * u, 'gradient') - 0.2 * math.laplace(x, padding=math.extrapolation.PERIODIC)


BOUNDARY_OPERATORS = {'neumann': neumann, 'periodic': periodic, 'mixed': mixed, 'varcoef': varcoef, 'periodic_advdiff': periodic_advdiff}
