"""Shared fixtures of the GPU test modules that use the Hubbard species-order handles (TEST INFRASTRUCTURE)."""
import os

import numpy as np
import pytest

import lin_builders as B
import species_builders as SB
import quantum_basis_b200 as qb
from quantum_basis_b200 import _lib

SPECIES = _lib.SPECIES_ORDER
TOL_MV, TOL_E0, TOL_KPM = 1e-12, 1e-10, 1e-9
CASES = {"hub4x2_35": (4, 2, 3, 5, 1.1), "hub4x2_44": (4, 2, 4, 4, 1.1), "hub3x3_45": (3, 3, 4, 5, 2.3), "hub2x2_12": (2, 2, 1, 2, 0.7)}


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _case(name):
    Lx, Ly, nu, nd, U = CASES[name]
    ns, bonds = Lx * Ly, B.square_bonds(Lx, Ly)
    mk = lambda cx=True, mf=False, flags=SPECIES: qb.hubbard(ns, nu, nd, bonds, 1.0, U, is_complex=cx, matrix_free=mf, flags=flags)   # noqa: E731
    return ns, nu, nd, bonds, U, mk
