/* TEST INFRASTRUCTURE ONLY: the six instantiations of the simulator kernel (scbonita_b200/csrc/scb_sim_tu.cu is built
 * once per instantiation by the product Makefile) in one host translation unit for the emulator build. */
#define SCB_SIM_K 1
#define SCB_SIM_TM 0
#include "scb_sim_tu.cu"
#undef SCB_SIM_K
#undef SCB_SIM_TM
#define SCB_SIM_K 1
#define SCB_SIM_TM 1
#include "scb_sim_tu.cu"
#undef SCB_SIM_K
#undef SCB_SIM_TM
#define SCB_SIM_K 2
#define SCB_SIM_TM 0
#include "scb_sim_tu.cu"
#undef SCB_SIM_K
#undef SCB_SIM_TM
#define SCB_SIM_K 2
#define SCB_SIM_TM 1
#include "scb_sim_tu.cu"
#undef SCB_SIM_K
#undef SCB_SIM_TM
#define SCB_SIM_K 4
#define SCB_SIM_TM 0
#include "scb_sim_tu.cu"
#undef SCB_SIM_K
#undef SCB_SIM_TM
#define SCB_SIM_K 4
#define SCB_SIM_TM 1
#include "scb_sim_tu.cu"
