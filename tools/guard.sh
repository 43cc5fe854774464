#!/bin/bash
# usage: tools/guard.sh <max_rss_gb> <timeout_s> cmd args...   -- runs cmd; kills it if its resident memory exceeds the
# limit or the timeout expires (a runaway host allocation must never take the GPU box down with it)
lim_kb=$(( $1 * 1024 * 1024 )); tmo=$2; shift 2
"$@" &
pid=$!
( end=$(( $(date +%s) + tmo ))
  while kill -0 $pid 2>/dev/null; do
    rss=$(awk '/VmRSS/{print $2}' /proc/$pid/status 2>/dev/null); rss=${rss:-0}
    if [ "$rss" -gt "$lim_kb" ]; then echo "guard: RSS ${rss} kB over the limit, killing $pid" >&2; kill -9 $pid; break; fi
    if [ "$(date +%s)" -ge "$end" ]; then echo "guard: timeout, killing $pid" >&2; kill -9 $pid; break; fi
    sleep 0.05
  done ) &
wait $pid
