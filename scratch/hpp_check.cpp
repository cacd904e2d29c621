#include "../include/triple_accel.hpp"
int main() { return 0; }
