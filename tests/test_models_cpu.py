"""Host-side model checks of kernel control flow (no GPU): see tools/models/."""
import importlib.util
import os

from conftest import ROOT


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", "models", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_inner_loop_restructuring_visits_the_same_ticks():
    """SF_INNER_LOOP (build option of csrc/sf_jacobi.cu) must not change which rows run which tick."""
    _load("loop_equivalence").main(trials=3000, seed=7)


def test_stream_kernel_model_matches_the_oracle_on_the_smallest_grids():
    """The numpy transcription of jacobi_stream_kernel (pipeline, fused set_bnd, fast groups, chunking, launch plan, implicit
    zero guess; unloaded data = NaN) is bit-identical to the oracle at G = 4, 8, 12, 32 and at a two-band width."""
    _load("stream_model").main(sizes=(2, 6, 10, 30, 114))


def test_temporally_blocked_red_black_design_matches_the_in_place_scheme():
    """Design check for the next step of the opt-in solver: red-black Gauss-Seidel / SOR on the streaming pipeline
    (one iteration = two levels, colour-masked updates, set_bnd on black levels) is bit-identical to the in-place scheme."""
    _load("rbgs_blocked_model").main(sizes=(2, 6, 14, 114))


def test_tma_staged_advect_model_matches_the_oracle():
    """numpy transcription of advect_tile_kernel (bounding box of the traces, first column rounded down to the TMA unit's 16-byte
    rule, fit test, 8-row tensor copies with zero fill, shared-memory index arithmetic, whole-tile fallback, 2^23 truncation,
    fused set_bnd): gathers of fitted tiles are served from the modelled box only, and the result is bit-identical to the
    oracle's advect for smooth, wall-hitting and partly rough velocity fields, both tile shapes, b = 0, 1, 2."""
    _load("advect_tile_model").main(sizes=(318, 382))
