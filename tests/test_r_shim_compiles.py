"""The R `.Call()` shim cannot run here (no R in the image); it is type-checked against a minimal
mock of Rinternals.h so that a signature drift between the shim and include/infercnv_b200.h fails CI."""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_r_shim_type_checks_against_the_c_abi():
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration",
                        "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
                        "-I", os.path.join(ROOT, "infercnv_b200", "r", "mock"), "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "infercnv_b200", "r", "infercnvb200_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_r_wrappers_name_every_replaced_function():
    src = open(os.path.join(ROOT, "infercnv_b200", "r", "infercnv_b200.R")).read()
    for fn in ["subtract_ref_expr_from_obs", "smooth_by_chromosome", "center_cell_expr_across_chromosome",
               "predict_CNV_via_HMM_on_indiv_cells", "predict_CNV_via_HMM_on_tumor_subclusters",
               "predict_CNV_via_HMM_on_whole_tumor_samples", "i3HMM_predict_CNV_via_HMM_on_indiv_cells",
               "apply_median_filtering", "get_predicted_CNV_regions", "remove_outliers_norm", "clear_noise",
               "predict_CNV_via_HMM_on_tumor_subclusters_per_chr", "i3HMM_predict_CNV_via_HMM_on_tumor_subclusters",
               "i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples", "normalize_counts_by_seq_depth",
               "clear_noise_via_ref_mean_sd"]:
        assert f"b200_{fn} <- function" in src and f'"{fn}"' in src
    # the one replaced IMPORT (parallelDist::parallelDist, R/inferCNV_constants.R:27): same signature, bound in the imports environment
    assert 'b200_parallelDist <- function(x, method = "euclidean", diag = FALSE, upper = FALSE, threads = NULL, ...)' in src
    assert 'assign("parallelDist", b200_parallelDist, envir = imp)' in src and "parent.env(ns)" in src


def test_r_wrappers_have_balanced_brackets_and_registered_call_names():
    """R is not installable here, so the R file gets the checks that do not need an interpreter: brackets balance
    outside strings and comments, and every .Call("icnvR_...") names a routine the shim registers."""
    import re
    src = open(os.path.join(ROOT, "infercnv_b200", "r", "infercnv_b200.R")).read()
    stack, line, in_str, i = [], 1, None, 0
    pairs = {")": "(", "]": "[", "}": "{"}
    while i < len(src):
        ch = src[i]
        if ch == "\n":
            line += 1
        if in_str:
            if ch == "\\":
                i += 1
            elif ch == in_str:
                in_str = None
        elif ch in "\"'":
            in_str = ch
        elif ch == "#":
            while i < len(src) and src[i] != "\n":
                i += 1
            line += 1
        elif ch in "([{":
            stack.append((ch, line))
        elif ch in ")]}":
            assert stack and stack[-1][0] == pairs[ch], f"unbalanced {ch!r} at line {line}"
            stack.pop()
        i += 1
    assert not stack and in_str is None, f"unclosed {stack[-1] if stack else in_str}"
    shim = open(os.path.join(ROOT, "infercnv_b200", "r", "infercnvb200_shim.c")).read()
    registered = dict(re.findall(r'\{"(icnvR_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', shim))
    calls = []
    for m in re.finditer(r'\.Call\("(icnvR_\w+)"', src):
        depth, n, j = 1, 0, m.end()
        while depth:                      # walk to the matching parenthesis, counting top-level commas
            ch = src[j]
            depth += ch in "([{"
            depth -= ch in ")]}"
            n += ch == "," and depth == 1
            j += 1
        calls.append((m.group(1), n))
    assert len(calls) >= 13
    for name, n in calls:
        assert name in registered, f"{name} is not registered in the shim"
        assert n == int(registered[name]), f"{name}: {n} arguments passed, {registered[name]} registered"
