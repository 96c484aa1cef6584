#!/bin/bash
# development aid: freeze a consistent copy of the tree under _snap/ so that a queued gpurun call (which snapshots
# /root/repo only when a GPU box is acquired) runs exactly what was built now, while editing continues in the main tree
cd "$(dirname "$0")/.." || exit 1
rm -rf _snap && mkdir _snap
tar cf - --exclude=./.git --exclude=./gpurun_out --exclude=./_snap --exclude=./_ab_old --exclude=.pytest_cache --exclude=__pycache__ . | tar xf - -C _snap
echo "snapshot: $(du -sh _snap | cut -f1)"
