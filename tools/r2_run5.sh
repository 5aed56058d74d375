#!/bin/bash
python tools/prof_cell.py 10000000 6
python tools/prof_cell.py 10000000 6 q7
python tools/prof_cell.py 10000000 6
python tools/prof_cell.py 10000000 6 q7
