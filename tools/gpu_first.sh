#!/bin/bash
# first GPU contact: parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
