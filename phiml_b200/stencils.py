"""
Constant-coefficient stencils of the operators the BASELINE configs use, as (offsets, coefficients) for
``Matrix.stencil_matrix`` — the workload definitions of bench.py and scripts/, product side (nothing here touches oracle/).

* ``neg_laplace``: ``-math.laplace(x, padding=ZERO)`` on a unit grid: diagonal 2·rank, the 2·rank neighbours −1 (configs 1–4).
* ``advection_diffusion``: ``x + c·(u·∇x) − ν·laplace(x)`` with central differences, ZERO padding (config 5; nonsymmetric).
  Coefficients are evaluated in the matrix dtype in the order PhiML's tracer evaluates them, so the fp32 values equal the ones in
  a matrix assembled by ``matrix_from_function`` bit for bit (tests/test_abi.py compares with the oracle's table).
"""
from typing import List, Sequence, Tuple

import numpy as np

Stencil = Tuple[List[Tuple[int, ...]], List[float]]


def _unit(rank: int, axis: int, step: int) -> Tuple[int, ...]:
    return tuple(step if a == axis else 0 for a in range(rank))


def neg_laplace(rank: int) -> Stencil:
    offsets, coeffs = [], []
    for axis in range(rank):
        for step in (-1, +1):
            offsets.append(_unit(rank, axis, step))
            coeffs.append(-1.0)
        if axis == 0:                      # the tracer meets the centre point after the first axis' neighbours
            offsets.append(tuple([0] * rank))
            coeffs.append(2.0 * rank)
    return offsets, coeffs


def advection_diffusion(rank: int, velocity: Sequence[float], c: float, nu: float, dtype=np.float32) -> Stencil:
    t = np.dtype(dtype).type
    offsets, coeffs = [tuple([0] * rank)], [float(t(1) - t(nu) * t(-2.0 * rank))]
    for axis in range(rank):
        transport = t(c) * (t(0.5) * t(velocity[axis]))
        offsets.append(_unit(rank, axis, +1))
        coeffs.append(float(transport - t(nu)))
        offsets.append(_unit(rank, axis, -1))
        coeffs.append(float(-transport - t(nu)))
    return offsets, coeffs


ADVDIFF = dict(velocity=(1.0, 0.5, 0.25), c=0.3, nu=0.2)      # BASELINE config 5


def neg_laplace_pairs(rank: int):
    """[(offset, coefficient), ...] form of ``neg_laplace``."""
    return list(zip(*neg_laplace(rank)))


def advection_diffusion_pairs(rank: int, velocity: Sequence[float], c: float, nu: float, dtype=np.float32):
    return list(zip(*advection_diffusion(rank, velocity, c, nu, dtype)))
