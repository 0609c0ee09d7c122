#!/bin/bash
timeout 900 python -m pytest tests/test_path.py -q -m gpu -x -k "queue_drain_order or interleaved or multi_sample" 2>&1 | tail -5
