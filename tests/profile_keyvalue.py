"""KeyValue head under ncu (launch list / --set full of kv_gemm_pair_kernel): a few calls at the 480p shapes.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/<tag>/kv_launches.csv \
        python tests/profile_keyvalue.py
Diagnostic script, not a test."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfloodnet_b200 as vfn  # noqa: E402


def main():
    torch.manual_seed(0)
    key = torch.nn.Conv2d(1024, 128, 3, padding=1).cuda()
    val = torch.nn.Conv2d(1024, 512, 3, padding=1).cuda()
    head = vfn.KeyValueHead(key, val, passes=int(os.environ.get('KV_PASSES', '3'))).eval()
    reps = int(os.environ.get('KV_REPS', '3'))
    for b in (1, 2):
        x = torch.relu(torch.randn(b, 1024, 30, 54, device='cuda')) * 3
        with torch.no_grad():
            for _ in range(reps):
                head(x)
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
