#!/bin/bash
# sweep lanes-per-read (G) and min-blocks-per-SM of the seeding kernel on a 20 Mbp genome
for cfg in "8 2" "8 3" "8 4" "16 2" "16 3" "16 4" "32 3" "32 4"; do
  set -- $cfg
  echo "== G=$1 MINB=$2"
  MEME_SEED_GROUP=$1 MEME_SEED_MINB=$2 G=${GENOME:-20000000} N=${NREADS:-400000} python profiles/prof_seed_small.py 2>&1 | tail -2
done
