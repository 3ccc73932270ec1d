"""One-off check: the builder's device path at 70_otus size for the discrete-Gamma database (the GTR one is a GPU test,
tests/test_gpu_ref.py::test_builder_on_the_device_at_70_otus_size): prints ptucmp against the reference-built .ptu."""
import os, pathlib, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_ref as T
bind = os.path.join(ROOT, "hmmufotu_b200", "bin")
for db, extra in (("gg_70_otus_HKY85_dG", ["-V", "-k", "4"]),):
    tmp = pathlib.Path(tempfile.mkdtemp())
    ref_ptu, seqs, tree, anno, sm = T._export_builder_inputs(db, tmp)
    out = tmp / "out"
    out.mkdir()
    r = subprocess.run([os.path.join(bind, "hmmufotu-build"), seqs, tree, "-a", anno, "-n", "re70", "--dm", os.path.join(T.HERE, "golden", "build", "test_prior.dm"),
                        "--sm", sm] + extra, cwd=str(out), capture_output=True, text=True)
    print(db, "rc", r.returncode, r.stderr.strip().split("\n")[-2:])
    print(subprocess.run([os.path.join(bind, "hu_host_test"), "ptucmp", str(out / "re70.ptu"), ref_ptu], capture_output=True, text=True).stdout.strip())
    print(subprocess.run([os.path.join(bind, "hu_host_test"), "dgshow", str(out / "re70.ptu")], capture_output=True, text=True).stdout.split("\n")[0],
          "| reference:", subprocess.run([os.path.join(bind, "hu_host_test"), "dgshow", ref_ptu], capture_output=True, text=True).stdout.split("\n")[0])
