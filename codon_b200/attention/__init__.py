"""Drop-in for the reference ``attention`` package (CODON_X4/attention/)."""
