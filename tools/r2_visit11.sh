#!/bin/bash
set -u
echo "== overlapped"; SPECS="8:0 8:3 4:1" tools/strip_study.sh 8192
echo "== frame events (no overlap)"; SPECS="8:0 8:3 4:1" EXTRA=--frame-events tools/strip_study.sh 8192
