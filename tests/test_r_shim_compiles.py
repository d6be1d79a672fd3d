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
               "apply_median_filtering", "get_predicted_CNV_regions"]:
        assert f"b200_{fn} <- function" in src and f'"{fn}"' in src
