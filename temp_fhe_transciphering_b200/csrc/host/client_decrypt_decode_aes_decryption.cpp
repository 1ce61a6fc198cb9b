#include "client_decrypt_common.h"
int main(int argc, char **argv) { return decrypt_stage(argc, argv, "ciphertext_aes_download", "decoded_result_aes.txt"); }
