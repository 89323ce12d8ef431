// s4f_comm.h -- host interface of the peer-memory exchange layer (s4f_comm.cu)
#pragma once
#include <memory>
#include <vector>

#include "s4f_ctx.h"

struct S4fHaloPlan;
struct S4fGatherPlan;

int s4f_allgather_host(s4fgpu_ctx* c, const void* send, size_t bytes, void* recv);
int s4f_allgatherv_host(s4fgpu_ctx* c, const void* send, size_t bytes, std::vector<std::vector<char>>& recv);
int s4f_exchange_nbr_ints(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<std::vector<int>>& send,
                          std::vector<std::vector<int>>& recv);

int s4f_comm_setup(s4fgpu_ctx* c);          // reduction mailboxes; collective, after s4fgpu_comm_init
void s4f_comm_destroy(s4fgpu_ctx* c);

// collective over all ranks (ranks without neighbours pass empty lists).  sendCells: the local cells whose values go to
// neighbour 0, then neighbour 1, ...; the ghost slots are filled in the same order from ghostBase on.
int s4f_halo_plan_create(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<int>& nbrCount, const std::vector<int>& sendCells,
                         int maxComp, S4fHaloPlan** out);
// the two directions may carry different numbers of values (coarse GAMG levels: distinct cells on either side)
int s4f_halo_plan_create_asym(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<int>& sendCount, const std::vector<int>& recvCount,
                              const std::vector<int>& sendCells, int maxComp, S4fHaloPlan** out);
void s4f_halo_plan_destroy(S4fHaloPlan* P);
template <class T>
int s4f_halo_run(s4fgpu_ctx* c, S4fHaloPlan* P, T* field, int ld, int ncomp, int ghostBase);

int s4f_gather_plan_create(s4fgpu_ctx* c, const std::vector<int>& cntPerRank, S4fGatherPlan** out);
void s4f_gather_plan_destroy(S4fGatherPlan* P);
template <class T>
int s4f_gather_run(s4fgpu_ctx* c, S4fGatherPlan* P, const T* src, int lds, T* dst, int ldd, const int* act);
