"""Calibration sweep for the tile-selection model: every layer shape x N x forced tile width / pairing."""
import io, os, sys, contextlib, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bench_conv
res = {}
for n in (31, 32, 63, 126):
    for bn in ("", "256", "128", "64"):
        for mode in ("", "0", "1"):
            if bn == "" and mode != "": continue
            if bn != "" and mode == "": continue
            for k, v in (("FPB_GEMM_BN", bn), ("FPB_GEMM_MODE", mode)):
                if v: os.environ[k] = v
                else: os.environ.pop(k, None)
            sys.argv = ["bench_conv.py", str(n)]
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                try:
                    bench_conv.main()
                except Exception as e:
                    print("ERR", e)
            for line in buf.getvalue().splitlines():
                if " ms " in line:
                    name = line[:18].strip()
                    ms = float(line.split(" ms ")[0].split()[-1])
                    res.setdefault(f"{name}|n{n}", {})[f"bn{bn or 'auto'}_m{mode or 'auto'}"] = round(ms * 1e3, 1)
for k, v in res.items():
    print(k, json.dumps(v))
