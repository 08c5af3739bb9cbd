#!/usr/bin/env python
"""Root-level launcher kept for drop-in use of the reference's command lines:
    python eval_ali.py <the reference's flags>
The implementation lives in whisper_char_alignment_b200/cli/eval_ali.py."""
from whisper_char_alignment_b200.cli.eval_ali import main

if __name__ == "__main__":
    main()
