"""Drop-in check of the header API on the CPU box: the reference's example
models of the four BASELINE configs compile UNCHANGED against include/ (all 23
examples do: scripts/compile_examples.py). Needs /root/reference to read the
sources in place; skipped where it is not mounted (e.g. on the GPU box)."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))
import compile_examples  # noqa: E402

CONFIG_EXAMPLES = ["springs", "passive_growth", "epithelium", "branching"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(compile_examples.REFERENCE,
                                                    "examples")),
                    reason="reference sources not mounted")
def test_config_examples_compile_unchanged():
    results = compile_examples.compile_examples(CONFIG_EXAMPLES, workers=4)
    failed = [(name, errors) for name, ok, errors in results if not ok]
    assert not failed, failed
