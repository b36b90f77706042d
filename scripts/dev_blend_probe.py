import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, json
import bench
print(json.dumps(bench.blend_roofline(torch.device("cuda", 0), 6517.6, "measured")))
