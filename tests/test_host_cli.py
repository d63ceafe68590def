"""CPU-side checks of the host drop-in binaries (tiebrush_b200/host): they exist after build() wherever the reference
checkout is present, link the C-ABI library, and FAIL LOUDLY without a CUDA device (no CPU fallback behind the ABI)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tiebrush_b200", "host", "_build")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.parametrize("tool,args", [("tiebrush_gpu", ["-o", "{tmp}/o.bam", "{sam}"]), ("tiecov_gpu", ["-c", "{tmp}/c", "{sam}"])])
def test_host_tool_fails_loudly_without_cuda(tool, args, tmp_path):
    exe = os.path.join(HOST, tool)
    if not os.path.exists(exe):
        pytest.skip("host binaries are built only where /root/reference exists")
    if not _no_gpu():
        pytest.skip("a CUDA device is present; the GPU suite covers the tool")
    samf = tmp_path / "a.sam"
    samf.write_text("@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:1000\nr1\t0\tchr1\t10\t60\t50M\t*\t0\t0\t*\t*\tNH:i:1\n")
    cmd = [exe] + [a.format(tmp=str(tmp_path), sam=str(samf)) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_host_tool_links_the_c_abi_library():
    exe = os.path.join(HOST, "tiebrush_gpu")
    if not os.path.exists(exe):
        pytest.skip("host binaries are built only where /root/reference exists")
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libtiebrush_b200.so" in out and "not found" not in out.split("libtiebrush_b200.so")[1].split("\n")[0]


def test_parallel_reader_contract_dry_run(tmp_path):
    """TB_DRYRUN: the parallel per-file reader of tiebrush_gpu (no device involved) must hand over every record once, in
    file order, in windows separated by coverage gaps — for one big window and for many small ones."""
    import re
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(HOST, "tiebrush_gpu")
    if not os.path.exists(exe):
        pytest.skip("host binaries are built only where /root/reference exists")
    import test_gpu_host_cli as T
    paths = T._write_sams(str(tmp_path), k=5, reads=3000, seed=5, n_tx=20)
    seen = set()
    for wr, sp, thr in ((1 << 20, 1 << 21, 4), (300, 2000, 3), (1, 200, 1)):
        r = subprocess.run([exe, "-o", str(tmp_path / "o.bam")] + paths, capture_output=True, text=True,
                           env=dict(os.environ, TB_DRYRUN="1", TB_WINDOW_RECORDS=str(wr), TB_WINDOW_SPAN=str(sp), TB_DECODE_THREADS=str(thr)))
        m = re.search(r"dry run: (\d+) windows, (\d+) records, (\d+) order violations, (\d+) gap violations", r.stderr)
        assert m, r.stderr[-300:]
        assert int(m.group(2)) == 5 * 3000 and int(m.group(3)) == 0 and int(m.group(4)) == 0
        seen.add(int(m.group(1)))
    assert min(seen) == 1 and max(seen) > 5
