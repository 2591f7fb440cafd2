"""tools/ncu_summary.py: the script that turns raw Nsight Compute CSV into the summaries committed under profiles/."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_list_summary(tmp_path):
    p = tmp_path / "l.csv"
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
    row = '"{i}","1","python","h","{k}","1","7","(288, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ns","{v}"\n'
    p.write_text("==PROF== noise\n" + hdr + row.format(i=0, k="void fq3::fq3_decode_kernel<1>(fq3::KParams)", v="3,000,000")
                 + row.format(i=1, k="void at::native::vectorized_elementwise_kernel<4>(int)", v="1000000")
                 + row.format(i=2, k="void fq3::fq3_decode_kernel<1>(fq3::KParams)", v="1000000"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches", str(p), "hdr"],
                         check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0] == "# hdr" and "3 launches, 5.0 ms total" in out[1] and "80.0%" in out[1]
    assert out[4].startswith("4000.0,80.00,2,fq3::fq3_decode_kernel<1>")


def test_kernel_summary(tmp_path):
    p = tmp_path / "r.csv"
    p.write_text('"ID","Kernel Name","Block Size","Grid Size","dram__bytes_read.sum","gpu__time_duration.sum","other"\n'
                 '"","","","","Gbyte","ms",""\n'
                 '"0","void fq3_decode_kernel<1>(KParams)","(288, 1, 1)","(148, 1, 1)","35.8","21.6","x"\n')
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "kernel", str(p), "hdr"],
                         check=True, capture_output=True, text=True).stdout.splitlines()
    assert "0,dram__bytes_read.sum,Gbyte,35.8" in out and "0,gpu__time_duration.sum,ms,21.6" in out
