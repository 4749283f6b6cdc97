"""One-GPU run of the pushed search step (score board in local memory): dot product with the gather fused in,
client wait, wire-form decrypt, credit.  Used under ncu for profiles/r1_ncu_push_step.txt.  usage: push_profile.py [docs] [steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import build_model, synthetic_docs, _decrypt_device
from fhe_icp_b200.score_board import PeerScoreBoard

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
m, _ = build_model(0)
_, _, X = synthetic_docs(docs, 5)
ct = m.encrypt(X)
board = PeerScoreBoard(m, docs)
post = torch.cuda.Stream(priority=-1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps + 3):
    if i == 3:
        torch.cuda.synchronize(); e0.record()
    board.push(ct)
    with torch.cuda.stream(post):
        slot = board.collect()
        _decrypt_device(m, slot, wire32=True)
        board.release()
torch.cuda.current_stream().wait_stream(post)
e1.record(); torch.cuda.synchronize()
board.check()
ok = np.array_equal(m.decrypt_compressed(slot[:docs]), m.predict_clear(X))
print(f"pushed step, {docs} documents on one GPU: {e0.elapsed_time(e1) / steps:.4f} ms per step, exact={ok}")
board.close()
