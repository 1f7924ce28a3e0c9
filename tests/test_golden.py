"""Committed golden fixtures (tests/golden/golden.json, SHA-256 of canonical output bytes, generated from the compiled reference by
tests/golden/make_golden.py): the plain-C oracle must reproduce them on the CPU, the CUDA path on the GPU -- also on a box where neither
/root/reference nor oracle/_ref exists."""
import json
import os

import pytest

import common
import golden_cases

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")) as f:
    GOLDEN = json.load(f)["sha256"]
CASES = golden_cases.cases()


def test_fixture_file_covers_every_case():
    assert sorted(GOLDEN) == sorted(k for k, _, _ in CASES)


@pytest.mark.parametrize("key,runner,case", CASES, ids=[k for k, _, _ in CASES])
def test_oracle_port_reproduces_golden(key, runner, case):
    assert runner(case, "port") == GOLDEN[key], f"{key}: oracle/port output differs from the committed reference digest"


@pytest.mark.skipif(common.load_ref() is None, reason="oracle/_ref not present")
@pytest.mark.parametrize("key,runner,case", CASES[::4], ids=[k for k, _, _ in CASES[::4]])
def test_compiled_reference_still_matches_golden(key, runner, case):
    """the fixtures are reproducible from the reference as it compiles today (a sample: every fourth case)"""
    assert runner(case, "reference") == GOLDEN[key]


CUDA_CASES = [c for c in CASES if not c[0].startswith(golden_cases.NO_CUDA_YET)]


@pytest.mark.gpu
@pytest.mark.parametrize("key,runner,case", CUDA_CASES, ids=[k for k, _, _ in CUDA_CASES])
def test_cuda_reproduces_golden(key, runner, case):
    assert runner(case, "cuda") == GOLDEN[key], f"{key}: CUDA output differs from the committed reference digest"
