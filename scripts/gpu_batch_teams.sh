timeout 600 python -m pytest tests -m gpu -q -x -k "batch or run_many" 2>&1 | tail -3
for n in 1024 128; do for b in 1 0; do if [ $b = 1 ]; then export SIGSDP_BATCH_BLOCKS=1; else unset SIGSDP_BATCH_BLOCKS; fi; python bench.py --workload cfg5_batch --instances $n --steps 150 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('instances',d['config']['instances'],'bpi',d['config']['blocks_per_instance'],'value %.0f'%d['value'],'ms/step %.3f'%d['ms_per_step'])"; done; done
