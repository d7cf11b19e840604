"""The reference's three command lines on the CUDA path (see tests/frontend_flow.py)."""
import pytest

from tests.frontend_flow import run_flow

pytestmark = pytest.mark.gpu


def test_train_record_eval_command_lines(tmp_path):
    run_flow(tmp_path, "cuda:0")
