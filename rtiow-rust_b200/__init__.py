"""rtiow-rust_b200: the B200-native render path of cbiffle/rtiow-rust.
Import as `rtiow_rust_b200` (the repo root carries an alias package for the hyphenated directory)."""
