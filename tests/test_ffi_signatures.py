"""Every FFI binding against include/abo.h: the `ccall` sites of julia/AboCuda.jl (the Julia shim cannot be executed in
this image, so its signatures are verified statically) and the ctypes table of the Python binding."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("check_ffi", os.path.join(ROOT, "tools", "check_ffi_signatures.py"))
chk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(chk)


def test_julia_ccalls_match_header():
    protos = chk.c_prototypes()
    n, errs = chk.check_julia(protos)
    assert n >= 25 and not errs, "\n".join(errs)
    bound = {c[0] for c in chk.julia_ccalls()}
    # the GradientGP surface and the fused acquisitions are bound, not only the StandardGP path
    for name in ("abo_gp_create", "abo_gp_fit", "abo_gp_append", "abo_gp_clone", "abo_gp_posterior", "abo_gp_posterior_cov",
                 "abo_acq_eval", "abo_acq_eval_grad", "abo_acq_eval_multi", "abo_nlml_batch", "abo_fill_distance",
                 "abo_gp_set_params", "abo_gp_set_params_ard", "abo_nlml_batch_ard", "abo_standardize"):
        assert name in bound, name


def test_ctypes_table_matches_header():
    protos = chk.c_prototypes()
    n, errs = chk.check_ctypes(protos)
    assert n >= 30 and not errs, "\n".join(errs)


def test_checker_catches_a_wrong_signature():
    protos = chk.c_prototypes()
    assert protos["abo_gp_fit"] == ("int32", ["ptr:void", "ptr:double", "ptr:double", "int64", "ptr:int64"])
    assert protos["abo_ctx_profile_read"][1] == ["ptr:void", "ptr:double", "ptr:int64"]          # arrays decay to pointers
    broken = dict(protos); broken["abo_gp_fit"] = ("int32", ["ptr:void", "ptr:double", "ptr:double", "int32", "ptr:int64"])
    _, errs = chk.check_julia(broken)
    assert any("abo_gp_fit" in e and "argument 4" in e for e in errs)
