"""CPU tests: the machine-specialised kernels generate and compile (NVRTC, sm_100a, no device)
for every small golden machine; big machines are reported as not eligible."""
import pytest

from helpers import FlatMachine, golden_names, load_golden


# one fixture per distinct machine structure (each check compiles nine kernels)
@pytest.mark.parametrize("name", ["bitnoise_tiny", "unitindel", "stutter_noise_difflen", "counter_xxx", "dnapsw_small", "protpsw_synth", "translate", "prot2dna_dnapsw", "dnapsw_dnapsw"])
def test_jit_kernels_compile(name):
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    log = capi.jit_compile_check(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    if fm.n_states > 16:      # mid-size machines: the big engine's generated thread-per-cell Forward sweep (mb_big.cu)
        assert "mb_k_big_forward" in log
        return
    assert "mb_k_forward" in log and "mb_k_viterbi" in log and "mb_k_backward" in log
    assert "0 bytes spill stores" in log


def test_split_mode_module_compiles():
    """The score module with MB_SPLIT 1 (strips of a pair as work items; built on the device only when a call with few,
    long pairs first needs it) compiles without spills; the ordinary modules carry no trace of it."""
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden("dnapsw_small")["machine"])
    capi.set_option("jit_split", 1)
    try:
        log = capi.jit_compile_check(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    finally:
        capi.set_option("jit_split", None)
    assert "split-mode module" in log
    tail = log[log.index("split-mode module"):]
    assert "mb_k_viterbi" in tail and "mb_k_forward_lin" in tail
    spills = [l for l in tail.splitlines() if "spill stores" in l]
    assert len(spills) == 4 and all(" 0 bytes spill stores" in l for l in spills), spills


@pytest.mark.parametrize("name,cols", [("dnapsw_small", 10), ("protpsw_synth", 10), ("dnapsw_small", 5), ("protpsw_synth", 12)])
def test_score_module_at_a_fitted_width_compiles(name, cols):
    """The score module at a strip width fitted to a batch (mb_jit.cu ensure_fit_module: MB_C not a power of two, the lane's
    Viterbi pointers in a padded 16-byte group, MB_TBPAD) is the same source as the wide module at that width: compile it."""
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden(name)["machine"])
    capi.set_option("jit_cv", cols)
    try:
        log = capi.jit_compile_check(fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    finally:
        capi.set_option("jit_cv", None)
    assert "MB_C = %d" % cols in log
    tail = log[log.index("Viterbi module"):]
    assert "mb_k_viterbi" in tail and "mb_k_forward_lin" in tail


def test_kernel_cache_directory(tmp_path):
    """mb_set_kernel_cache_dir: the modules compiled for a machine structure are kept as <hash of the source>.cubin, and the next
    compilation of the same structure -- in this or a later process -- takes them from there instead of running NVRTC."""
    import os
    import time
    from machineboss_b200 import capi
    fm = FlatMachine.from_json(load_golden("unitindel")["machine"])
    args = (fm.n_states, fm.n_in, fm.n_out, fm.src, fm.dst, fm.tin, fm.tout)
    capi.set_kernel_cache_dir(str(tmp_path))
    try:
        t0 = time.time()
        first = capi.jit_compile_check(*args)
        t1 = time.time()
        again = capi.jit_compile_check(*args)
        t2 = time.time()
    finally:
        capi.set_kernel_cache_dir(None)
    files = sorted(os.listdir(str(tmp_path)))
    assert len(files) >= 1 and all(f.endswith(".cubin") and len(f) == 22 for f in files), files
    assert "kernel cache" not in first and "kernel cache" in again
    assert (t2 - t1) < 0.5 * (t1 - t0)
    assert "kernel cache" not in capi.jit_compile_check(*args)      # switched off again: compiled
