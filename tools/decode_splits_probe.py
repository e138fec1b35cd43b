import sys, os
sys.path.insert(0, os.getcwd())
from tools.decode_trace import case
for S in (32, 36, 37, 24, 16):
    case(1, 32, 8, 32768, splits=S, target=5000)
for S in (4, 8):
    case(8, 32, 8, 8192, splits=S, target=6000)
for S in (8, 9):
    case(4, 32, 8, 16384, splits=S)
