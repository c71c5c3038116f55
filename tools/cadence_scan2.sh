for o in "flush_every=8 flush_threshold=16" "flush_every=8 flush_threshold=14" "flush_every=8 flush_threshold=12" "flush_every=6 flush_threshold=18" "flush_every=6 flush_threshold=14" "flush_every=10 flush_threshold=14" "flush_every=12 flush_threshold=12" "flush_every=8 flush_threshold=20" "flush_every=8 flush_threshold=24"; do
  args=""; for kv in $o; do args="$args --opt $kv"; done
  echo "== $o"; timeout 200 python tools/quick_bench.py --walkers 4096 --sweeps 864 --therm 432 $args 2>&1 | python -c "
import sys,json
t=sys.stdin.read(); i=t.index('{'); j=t.rindex('}')
d=json.loads(t[i:j+1]); tm=d['timers']; print(round(d['walker_sweeps_per_s']/1e6,2),'M/s', {k:round(v['ms'],1) for k,v in tm.items() if v['ms']>0}, 'flushes', tm['update']['flushes'])"
done
