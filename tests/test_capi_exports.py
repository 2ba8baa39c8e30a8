"""The C-ABI library loads and exports every symbol include/bpt_c_api.h declares (no compute calls)."""
import ctypes
import re

from tests.oracle_lib import REPO


def declared_functions():
    text = (REPO / "include" / "bpt_c_api.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ["bpt_create", "bpt_destroy", "bpt_upload_mesh", "bpt_set_instances", "bpt_build_accel", "bpt_set_materials",
                 "bpt_set_lights", "bpt_set_environment", "bpt_render", "bpt_bsdf_eval_sample_pdf", "bpt_intersect", "bpt_rng_sample4"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from bifrost3d_b200 import library_path
    lib = ctypes.CDLL(str(library_path()))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_lists_every_declared_symbol():
    from bifrost3d_b200.capi import EXPORTS
    assert sorted(EXPORTS) == declared_functions()


def test_no_cpu_fallback_without_a_device():
    import torch
    import pytest
    import bifrost3d_b200 as b
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(b.BptError):
        b.Bpt(0)
