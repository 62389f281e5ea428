for B in 1 2 4 8; do
  echo "B=$B tile: $(RC_FLOW_KERNEL=tile python tools/b1_profile.py $B 2>/dev/null)"
  for S in 12 16 24 32 48; do
    echo "B=$B strip seg=$S: $(RC_FLOW_KERNEL=strip RC_STRIP_SEG=$S python tools/b1_profile.py $B 2>/dev/null)"
  done
done
