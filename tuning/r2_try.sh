#!/bin/bash
bash tuning/r2_ab.sh "" sc0 sc70 sc80 sc90 h12 h10 sc0
