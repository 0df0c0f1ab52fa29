#!/bin/bash
cd /root/repo
timeout 35 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
