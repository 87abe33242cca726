run() { env "$@" python bench.py --workload c4 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$*', d['ms_per_step'],d['accept_rate'],d['swap_stage']['accepted_exchanges_per_call'], d['swap_stage']['ms_per_call'])"; }
run A=1
run SCORUS_PDL=0
run SCORUS_SWAP_DIRECT_STORE=1
run SCORUS_SWAP_DIRECT_STORE=1 SCORUS_PDL=0
run SCORUS_SWAP_FUSE=0 SCORUS_PDL=0
run SCORUS_SWAP_FUSE=0
run SCORUS_DYNAMIC_TILES=0
run SCORUS_DYNAMIC_TILES=0 SCORUS_PDL=0
