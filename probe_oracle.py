#!/usr/bin/env python
"""Root-level launcher kept for drop-in use of the reference's command lines:
    python probe_oracle.py <the reference's flags>
The implementation lives in whisper_char_alignment_b200/cli/probe_oracle.py."""
from whisper_char_alignment_b200.cli.probe_oracle import main

if __name__ == "__main__":
    main()
