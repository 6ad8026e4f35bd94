"""DPP weak forms as data (perphil.forms.dpp, forms/dpp.py:7-247).

The reference builds UFL forms; what the hot path needs from them is which blocks of
    A = (1/mu) [[k1 K + beta M, -beta M], [-beta M, k2 K + beta M]]
a form denotes.  A `DPPForm` records exactly that (4 integrals, rank 2 for the monolithic form --
the structure pinned by forms/_tests/test_dpp_regressions/test_dpp_form_structure_regression.yml),
so `get_matrix_data_from_form(a, bcs)` and `solve_dpp` accept it like the UFL `a`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

from .mesh import Function, MixedFunctionSpace
from .parameters import DPPParameters


@dataclass(frozen=True)
class Integral:
    kind: str          # "stiffness" | "mass"
    coefficient: float
    test_field: int
    trial_field: Optional[int]   # None: the field is a known (delayed) coefficient -> linear form


@dataclass(frozen=True)
class DPPForm:
    space: object
    params: DPPParameters
    rank: int
    blocks: Tuple[Tuple[int, int], ...]      # (row, col) blocks of A present in the form
    integrals_: Tuple[Integral, ...] = field(default_factory=tuple)
    coefficient_function: Optional[Function] = None

    def integrals(self):
        return self.integrals_

    def arguments(self):
        return tuple(range(self.rank))


def _check_mixed(W):
    if not hasattr(W, "num_sub_spaces") or W.num_sub_spaces() != 2:  # forms/dpp.py:113-114
        raise ValueError(f"Expected a 2-field MixedFunctionSpace, got {type(W)}")


def _scale_integrals(prm: DPPParameters, field_id: int, other_known: bool):
    k = float(prm.k1) if field_id == 0 else float(prm.k2)
    mu, beta = float(prm.mu), float(prm.beta)
    other = 1 - field_id
    # macro: (k1/mu) grad p1.grad q1 - xi q1 ; micro: (k2/mu) grad p2.grad q2 + xi q2 ; xi = -beta/mu (p1 - p2)
    return (
        Integral("stiffness", k / mu, field_id, field_id),
        Integral("mass", beta / mu, field_id, field_id),
        Integral("mass", -beta / mu, field_id, None if other_known else other),
    )


def dpp_form(W, model_params: DPPParameters):
    """Monolithic bilinear form and (zero) linear form (forms/dpp.py:95-132)."""
    _check_mixed(W)
    macro = _scale_integrals(model_params, 0, False)
    micro = _scale_integrals(model_params, 1, False)
    # UFL keeps the transfer term as one integral per scale: 2 + 2 integrals, rank 2
    ints = (macro[0], Integral("mass-transfer", macro[1].coefficient, 0, None), micro[0],
            Integral("mass-transfer", micro[1].coefficient, 1, None))
    a = DPPForm(W, model_params, 2, ((0, 0), (0, 1), (1, 0), (1, 1)), ints)
    L = DPPForm(W, model_params, 1, (), ())
    return a, L


def dpp_delayed_form(macro_function_space, micro_function_space, model_params: DPPParameters,
                     macro_pressure_initial_values: Function, micro_pressure_initial_values: Function):
    """Per-scale forms with the other pressure delayed (forms/dpp.py:135-205):
    a_i = (k_i/mu) grad p.grad q + (beta/mu) p q ;  L_i = (beta/mu) p_other_old q."""
    a_macro = DPPForm(macro_function_space, model_params, 2, ((0, 0),), _scale_integrals(model_params, 0, True)[:2])
    L_macro = DPPForm(macro_function_space, model_params, 1, (), (), micro_pressure_initial_values)
    a_micro = DPPForm(micro_function_space, model_params, 2, ((1, 1),), _scale_integrals(model_params, 1, True)[:2])
    L_micro = DPPForm(micro_function_space, model_params, 1, (), (), macro_pressure_initial_values)
    return (a_macro, L_macro), (a_micro, L_micro)


def dpp_splitted_form(W, model_params: DPPParameters):
    """Residual form + the Function holding (p1, p2) (forms/dpp.py:208-247)."""
    _check_mixed(W)
    fields = Function(W)
    F = DPPForm(W, model_params, 1, ((0, 0), (0, 1), (1, 0), (1, 1)), (), fields)
    return F, fields
