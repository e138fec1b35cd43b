"""Split-KV decode cases (few sequences, long context: the combine pass runs): time per call with the combine pass
launched programmatically behind the split kernel (default) and as a plain stream-ordered launch (PLI_NO_PDL=1):
    python tools/decode_split_perf.py ; PLI_NO_PDL=1 python tools/decode_split_perf.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.decode_sweep import run

if __name__ == "__main__":
    print("PDL", "off" if os.environ.get("PLI_NO_PDL") == "1" else "on")
    run(1, 32, 8, 32768, reps=50)
    run(1, 32, 8, 131072, reps=50)
    run(4, 32, 8, 16384, reps=50)
    run(8, 32, 8, 8192, reps=50)
    run(256, 4, 1, 1024, reps=50, splits=2)
