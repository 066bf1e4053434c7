"""CPU-only: the product library loads and exports every symbol include/esvo_b200.h declares
(no compute calls without a GPU), and esvo_create fails loudly without a device."""
import ctypes as C
import os
import re

from esvo_b200 import capi, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "esvo_b200.h")).read()
    return sorted(set(re.findall(r"ESVO_API\s+[\w\s\*]+?\b(esvo_\w+)\s*\(", src)))


def test_header_symbols_exported(product_lib, oracle_lib):
    names = _declared()
    assert len(names) >= 33
    for n in names:
        assert hasattr(product_lib.lib, n), f"product library does not export {n}"
    # the oracle mirrors the compute entry points under its own prefix
    for n in names:
        # device-side plumbing has no CPU counterpart
        if n.endswith("_dev") or n in ("esvo_stage_ts_events", "esvo_run_ts_build", "esvo_stage_mapping_inputs",
                                       "esvo_run_mapping", "esvo_fetch_mapping_counters", "esvo_sync", "esvo_stream",
                                       "esvo_launch_count", "esvo_last_error", "esvo_debug_counter", "esvo_profile",
                                       "esvo_profile_read", "esvo_set_pipeline_depth", "esvo_results_begin", "esvo_results_end", "esvo_results_end_view",
                                       "esvo_compute_rectify_tables",   # host set-up of the product, pinned to cv2 directly
                                       "esvo_sgbm_compute"):   # its oracle is oracle/sgbm.py (numpy, pinned against cv2)
            continue
        assert hasattr(oracle_lib.lib, n.replace("esvo_", "esvo_oracle_", 1)), n


def test_struct_layouts_match_header(product_lib):
    assert C.sizeof(capi.Calib) == 16 + 8 * (9 + 4 + 9 + 12)
    p = capi.default_params(product_lib)
    assert p.decay_ms == 30 and p.max_event_queue_len == 20 and p.num_thread_mapping == 4
    assert p.trk_batch_size == 200 and abs(p.bm_zncc_threshold - 0.1) < 1e-15 and p.max_iteration == 10


def test_create_fails_without_device(product_lib):
    import torch
    if torch.cuda.is_available():
        return
    l, r = configs.rig_calibs("hkust")
    try:
        capi.Backend(product_lib, l, r, configs.params_for("hkust", product_lib))
    except capi.EsvoError as e:
        assert e.code == -2
    else:
        raise AssertionError("esvo_create must fail without a CUDA device (no CPU fallback)")


def test_header_is_plain_c(tmp_path):
    """include/esvo_b200.h is a C ABI: it must compile as C99 (cgo / JNI / ctypes-style bindings include it from C)."""
    import subprocess
    src = tmp_path / "cabi.c"
    src.write_text('#include "esvo_b200.h"\nint main(void){ esvo_params p; esvo_default_params(&p); '
                   'return (int)sizeof(esvo_seed) + (int)sizeof(esvo_depth_point) == 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                           "-fsyntax-only", str(src)])
