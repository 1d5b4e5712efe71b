/* placeholder translation unit; the CPU folding oracle lands here */
int trxo_fold_abi_version(void) { return 0; }
