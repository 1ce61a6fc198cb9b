#include "client_decrypt_common.h"
int main(int argc, char **argv) { return decrypt_stage(argc, argv, "ciphertexts_download", "decoded_result.txt"); }
