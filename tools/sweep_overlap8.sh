run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NPROC:-8} --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/diag_overlap.py 2>&1 | grep -E "world|gram kernel|Error|error" ; }
run SQFA_GRAM_OVERLAP=0
run SQFA_GRAM_OVERLAP=1 SQFA_GRAM_RESERVE_SMS=8 SQFA_GRAM_GROUPS=5
run SQFA_GRAM_OVERLAP=1 SQFA_GRAM_RESERVE_SMS=8 SQFA_GRAM_GROUPS=10
run SQFA_GRAM_OVERLAP=1 SQFA_GRAM_RESERVE_SMS=16 SQFA_GRAM_GROUPS=10
