"""Static checks of the Julia shim (quantumaces.jl_b200/julia/ACESB200.jl).  Julia cannot run in this image,
so what can be checked is checked: every ccall against the C header (tools/check_ccall.py), block
structure, that install!() overrides the functions the integration document names, and that no name
QuantumACES exports is exported again."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import check_ccall  # noqa: E402

SHIM = os.path.join(ROOT, "quantumaces.jl_b200", "julia", "ACESB200.jl")


def test_every_ccall_matches_the_header():
    protos, bound, errors = check_ccall.check()
    assert not errors, "\n".join(errors)
    # the whole hot path is bound
    for name in ("aces_init", "aces_tables_create", "aces_parity_counts", "aces_parity_counts_rowmajor",
                 "aces_estimate_experiment", "aces_estimate_gate_eigenvalues", "aces_design_create", "aces_design_fit",
                 "aces_calc_covariance", "aces_design_inv_factor", "aces_design_ls_covariance",
                 "aces_design_precision_blocks", "aces_project_simplex", "aces_project_nonnegative_batched",
                 "aces_host_alloc"):
        assert name in bound, name
    assert len(check_ccall.parse_ccalls()) >= 35


def test_checker_detects_a_wrong_type(tmp_path):
    text = open(SHIM).read()
    bad = text.replace("(Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64})",
                       "(Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64})")
    assert bad != text
    p = tmp_path / "bad.jl"
    p.write_text(bad)
    protos = check_ccall.parse_header()
    errs = 0
    for name, ret, tlist, values, line in check_ccall.parse_ccalls(str(p)):
        cret, cargs = protos[name]
        errs += len(tlist) != len(cargs) or any(not check_ccall.type_ok(j, c) for j, c in zip(tlist, cargs))
    assert errs == 1


def _strip(text):
    text = re.sub(r'"""(?:.|\n)*?"""', '""', text)
    text = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', text)
    return re.sub(r"#[^\n]*", "", text)


def test_block_structure_balances():
    text = _strip(open(SHIM).read())
    openers = {"function", "if", "for", "while", "begin", "let", "try", "struct", "module", "quote", "macro", "do"}
    depth_paren, stack = 0, []
    for m in re.finditer(r"[()\[\]{}]|\b[A-Za-z_!]\w*\b", text):
        tok = m.group(0)
        if tok in "([{":
            depth_paren += 1
        elif tok in ")]}":
            depth_paren -= 1
            assert depth_paren >= 0, f"unbalanced bracket near offset {m.start()}"
        elif depth_paren == 0:
            if tok in openers:
                prev = text[max(0, m.start() - 8):m.start()]
                if tok == "struct" and prev.rstrip().endswith("mutable"):
                    pass
                stack.append((tok, text.count("\n", 0, m.start()) + 1))
            elif tok == "end":
                assert stack, f"`end` without an opener at line {text.count(chr(10), 0, m.start()) + 1}"
                stack.pop()
    assert depth_paren == 0
    assert not stack, f"unclosed blocks: {stack[:5]}"


def test_install_overrides_the_documented_functions():
    text = open(SHIM).read()
    body = text[text.index("function install!()"):]
    for name in ("simulate_stim_estimate", "estimate_experiment", "estimate_gate_noise", "estimate_gate_eigenvalues",
                 "calc_covariance", "calc_gls_covariance", "calc_wls_covariance", "calc_ols_covariance"):
        assert re.search(rf"^\s+{name}\(", body, flags=re.M), name
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("simulate_stim_estimate", "estimate_gate_noise", "calc_gls_covariance", "check_ccall.py", "abi_host.c"):
        assert name in doc, f"INTEGRATION.md does not mention {name}"


def test_no_export_clashes_with_quantumaces():
    text = _strip(open(SHIM).read())
    exported = re.search(r"^export (.*(?:\n    .*)*)", text, flags=re.M).group(1)
    names = {n.strip() for n in exported.replace("\n", " ").split(",")}
    ref_exports = os.path.join("/root/reference", "src", "QuantumACES.jl")
    clash = {"calc_covariance", "sparse_covariance_inv_factor", "estimate_gate_noise", "simulate_stim_estimate",
             "estimate_gate_eigenvalues", "estimate_experiment", "calc_gls_covariance"}
    if os.path.exists(ref_exports):   # this container only: the GPU box has no /root/reference
        ref = open(ref_exports).read()
        clash |= {n for n in names if re.search(rf"\b{re.escape(n)}\b", ref.split("# Public functions")[-1]) and
                  re.search(rf"^\s+{re.escape(n)},?$", ref, flags=re.M)}
    assert not (names & clash), names & clash
