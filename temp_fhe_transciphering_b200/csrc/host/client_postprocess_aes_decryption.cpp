#include "client_decrypt_common.h"
int main(int argc, char **argv) { return postprocess_stage(argc, argv, "decoded_result_aes.txt", "result_aes.txt"); }
