#!/bin/bash
# round-2 GPU job 3E: the seeded client kernel with the tail of its ciphertexts handed out dynamically (ticket counter) -- parity tests, then the e2e call A/B
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q --timeout 120 2>&1 | tail -3
{
echo "== static (FHE_B200_ENC_STATIC=1)"; FHE_B200_ENC_STATIC=1 timeout 120 python tools/e2e_ab.py 1000 2>&1 | tail -3
echo "== dynamic, 5/8 static (shipped build)"; timeout 120 python tools/e2e_ab.py 1000 2>&1 | tail -3
for v in es0 es3 es7; do echo "== dynamic variant $v"; FHE_B200_LIB=build_ab/libfhe_$v.so timeout 120 python tools/e2e_ab.py 1000 2>&1 | tail -2; done
echo "== static again"; FHE_B200_ENC_STATIC=1 timeout 120 python tools/e2e_ab.py 1000 2>&1 | tail -2
} > $O/r3e_enc_ticket_ab.txt 2>&1; cat $O/r3e_enc_ticket_ab.txt
