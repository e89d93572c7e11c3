"""Runs the BODIES of GPU parity tests against a NumPy-backed stand-in for the library handle
(tests/standin.py) on a machine without a GPU.  What this proves: the GPU tests themselves are sound -- their
fixture keys, record columns, state plumbing and tolerances are met by an independent correct implementation --
so that a failure on the B200 points at the CUDA path, not at the test.  It says nothing about the product."""
import numpy as np
import pytest

import standin
import test_golden as tg
import test_gpu_zstatus as tz


@pytest.fixture
def fake(fos, monkeypatch):
    monkeypatch.setattr(tg, "load_conic", standin.load_conic)
    monkeypatch.setattr(tz, "load_conic", standin.load_conic)
    import helpers
    monkeypatch.setattr(helpers, "load_affine", standin.load_affine)
    return standin.standin_module(fos)


@pytest.mark.parametrize("name", tg.FIXTURES)
def test_body_of_gpu_lockstep_on_golden(fake, name):
    tg.test_gpu_lockstep_on_golden(fake, name)


@pytest.mark.parametrize("name", tg.FEAS_FIXTURES)
def test_body_of_gpu_lockstep_on_feasibility_golden(fake, name):
    tg.test_gpu_lockstep_on_feasibility_golden(fake, name)


@pytest.mark.parametrize("kind,alg", [("infeasible", "DR"), ("infeasible", "Dykstra"), ("unbounded", "DR"),
                                      ("unbounded", "GAPA"), ("unbounded", "GAP")])
def test_body_of_status_branches_lockstep(fake, oracle, kind, alg):
    tz.test_status_branches_lockstep(fake, oracle, kind, alg)


@pytest.mark.parametrize("kind,alg", [("infeasible", "DR"), ("unbounded", "DR"), ("unbounded", "GAPA")])
def test_body_of_status_branches_free_running(fake, oracle, kind, alg):
    tz.test_status_branches_free_running(fake, oracle, kind, alg)


def test_body_of_cg_iteration_cap(fake, oracle):
    tz.test_cg_iteration_cap_sets_the_warning(fake, oracle)


@pytest.mark.parametrize("alg", ["DR", "GAPA", "Dykstra"])
def test_body_of_status_branches_batch_mode(fake, oracle, alg):
    tz.test_status_branches_batch_mode(fake, oracle, alg)
