for U in 15 20 24 30 36 45; do echo "units=$U"; TG_SCAN_UNITS=$U python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --variants 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(' literal kernel_ms %.4f frac %.3f | full kernel_ms %.4f frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['variants']['full_numeric_set']['kernel_ms'], d['variants']['full_numeric_set']['frac']))"; done
