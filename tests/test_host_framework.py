"""The C++ host side: the reference's plugin surface (Dwarf / Meter / Result / Registry / RunOptions /
makeMeasurements / CLI) re-created over the C ABI.  CPU part: formats, parsing, registry, error behaviour,
exported symbols.  GPU part: the reference's dwarf_tests (every dwarf x sizes x 10 iterations, all valid)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dwarf_bench_b200", "lib")


@pytest.fixture(scope="module")
def built():
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    # always through make (it knows what is up to date): a stale library must never be what gets tested
    subprocess.run(["make", "-C", os.path.join(ROOT, "dwarf_bench_b200", "csrc")], check=True, env=env, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", os.path.join(ROOT, "dwarf_bench_b200", "host")], check=True, env=env, stdout=subprocess.DEVNULL)
    return LIB


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def test_c_abi_exports_every_declared_symbol(built):
    """include/dwj.h is the contract: every DWJ_API declaration is an exported, loadable symbol -- and nothing else is."""
    header = open(os.path.join(ROOT, "include", "dwj.h")).read()
    declared = set(re.findall(r"DWJ_API [\w \*]*?\b(dwj_\w+)\(", header))
    assert len(declared) >= 14
    nm = run(["nm", "-D", "--defined-only", os.path.join(built, "libdwj_b200.so")]).stdout
    exported = set(re.findall(r" T (dwj_\w+)", nm))
    assert exported == declared
    from dwarf_bench_b200 import capi
    assert set(capi.SYMBOLS) == declared
    lib = capi.load_library()                       # loads without a GPU; only creating an engine needs one
    assert lib.dwj_abi_version() == 2
    assert capi.partition_of(12345, 4, 8) in range(8) and capi.partition_of(12345, 8, 1) == 0


def test_ctypes_binding_matches_header(built):
    """The ctypes binding (dwarf_bench_b200/capi.py) declares, for every entry point of include/dwj.h, exactly as many
    arguments as the prototype has -- a drifted binding would corrupt the stack silently."""
    from dwarf_bench_b200 import capi
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dwj.h")).read(), flags=re.S)
    protos = dict(re.findall(r"DWJ_API [\w \*]*?\b(dwj_\w+)\(([^;]*?)\);", header, flags=re.S))
    assert set(protos) == set(capi.SYMBOLS)
    lib = capi.load_library()
    for name, args in protos.items():
        args = " ".join(args.split())
        n_args = 0 if args in ("", "void") else args.count(",") + 1
        argtypes = getattr(lib, name).argtypes
        if n_args == 0:
            assert not argtypes, name
        else:
            assert argtypes is not None and len(argtypes) == n_args, (name, n_args, argtypes)


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dwarf_bench_b200 as dwj
    with pytest.raises(dwj.DwjError) as ei:
        dwj.Engine(100)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_host_tests_cpu(built):
    r = run([os.path.join(built, "host_tests"), "cpu"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


def test_cli_surface(built):
    cli = os.path.join(built, "dwarf_bench")
    r = run([cli, "list"])
    assert r.returncode == 0
    assert "DWARF_BENCH_ROOT is set to" in r.stdout and "Supported dwarfs:" in r.stdout
    listed = re.findall(r"^\t(\w+)$", r.stdout, flags=re.M)
    assert listed == ["CuckooHashBuild", "GroupBy", "GroupByCuda", "HashBuild", "HashBuildNonBitmask", "Join", "JoinOmnisci", "JoinOmnisciCuda", "SlabHashBuild",
                      "SlabJoin", "SlabProbe"]
    r = run([cli, "Join", "--help"])
    assert r.returncode == 0 and "--input_size arg" in r.stdout and "--report_path arg" in r.stdout
    r = run([cli, "NoSuchDwarf"])
    assert r.returncode == 1 and "List supported dwarfs to run with" in r.stderr
    r = run([cli, "Join", "--device=CPU", "--input_size", "128", "256"])       # case-insensitive device, multitoken
    assert r.returncode == 2 and "Caught exception: Join: this build serves --device=gpu only" in r.stderr
    r = run([cli, "Join", "--input_size=abc"])
    assert r.returncode == 2 and "Caught exception" in r.stderr
    r = run([cli, "Join", "--bogus=1"])
    assert r.returncode == 2 and "unrecognised option" in r.stderr


@pytest.mark.gpu
def test_host_tests_gpu(built):
    r = run([os.path.join(built, "host_tests"), "gpu"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_cli_join_on_gpu(built, tmp_path):
    """`dwarf_bench Join --device=gpu --input_size=... --iterations=...`: stdout block and CSV as the reference writes them."""
    cli = os.path.join(built, "dwarf_bench")
    r = run([cli, "Join", "--device=gpu", "--input_size", "1024", "65536", "--iterations=3"], env={**os.environ, "DWARF_BENCH_SEED": "5"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Selected device:" in r.stdout and "Incorrect results" not in r.stderr
    blocks = re.findall(r"Kernel duration: [\d.e+-]+ us\nHost duration:   [\d.e+-]+ us\nBuild time: [\d.e+-]+ us\nProbe time: [\d.e+-]+ us\n", r.stdout)
    assert len(blocks) == 6                                     # 2 sizes x 3 iterations
    report = tmp_path / "report.csv"
    for _ in range(2):                                          # append semantics: header once
        r = run([cli, "Join", "--device", "gpu", "--input_size=4096", "--iterations", "2", f"--report_path={report}"])
        assert r.returncode == 0, r.stdout + r.stderr
    lines = report.read_text().splitlines()
    assert lines[0] == "device_type,buf_size_bytes,host_time_ms,kernel_time_ms,build_time_ms,probe_time_ms"
    assert len(lines) == 5
    for row in lines[1:]:
        cells = row.split(",")
        assert cells[0] == "GPU" and cells[1] == str(4096 * 4) and len(cells) == 6 and all(float(c) >= 0 for c in cells[2:])
    for name in ("HashBuild", "SlabProbe", "JoinOmnisciCuda", "SlabJoin", "HashBuildNonBitmask", "SlabHashBuild", "CuckooHashBuild"):
        r = run([cli, name, "--device=gpu", "--input_size=2048", "--iterations=2"])
        assert r.returncode == 0 and r.stdout.count("Host duration:") == 2, name + r.stdout + r.stderr


@pytest.mark.gpu
def test_library_facade_bench_usage(built):
    """example/bench_usage: makeMeasurements{Join, GPU} returns one Measurement per iteration (the reference asserts here)."""
    r = run([os.path.join(built, "bench_usage")])
    assert r.returncode == 0, r.stdout + r.stderr
    assert len(re.findall(r"^1 1 RESULT: 1024 \d+$", r.stdout, flags=re.M)) == 10


@pytest.mark.gpu
def test_dropin_against_reference_framework():
    """oracle/_ref/dropin_join_test = our dwarfs compiled against the REFERENCE'S headers and linked with its own
    result/meter/options/registry sources (built only where the reference checkout exists; the binary travels)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_join_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_join_test not built (needs the reference checkout at build time)")
    r = run([exe])
    assert r.returncode == 0 and "dropin: 0 failures" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_join_over_two_gpus(built):
    """`dwarf_bench Join --device=gpu --gpus 2`: the C++ host reaches the multi-GPU join (dwj_mg_join_host) and every
    iteration matches the host's expected rows (no "Incorrect results", exit code 0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    env = dict(os.environ, DWARF_BENCH_SEED="11", DWJ_XJ_TIMEOUT_MS="10000")
    r = run([os.path.join(built, "dwarf_bench"), "Join", "--device=gpu", "--gpus", "2", "--input_size", "4096", "300000", "--iterations", "3"], env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Incorrect results" not in r.stderr and r.stdout.count("Build time:") == 6
