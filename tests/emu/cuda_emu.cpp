// TEST INFRASTRUCTURE ONLY -- runtime of the SIMT emulator declared in cuda_emu.h (fiber scheduler).
#include "cuda_emu.h"

thread_local EmuBlock *emu_blk = nullptr;
long emu_launches = 0;
static const size_t EMU_STACK = 256 * 1024;

char *emu_dyn_smem_ptr() { return emu_blk->dyn_smem; }

void emu_yield()
{
    EmuBlock *b = emu_blk;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

void emu_barrier(int group, int group_size)
{
    EmuBlock *b = emu_blk;
    if (group_size <= 1) return;
    const unsigned g = b->gen[group];
    if (++b->arrive[group] == group_size) {
        b->arrive[group] = 0;
        b->gen[group] = g + 1;
        return;
    }
    while (b->gen[group] == g) emu_yield();
}

static void fiber_entry()
{
    EmuBlock *b = emu_blk;
    b->body();
    b->fibers[b->cur].done = true;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

void emu_run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    emu_launches++;
    const int nthreads = (int)block.x;
    EmuBlock blk;
    blk.fibers.resize(nthreads);
    blk.bdim = block;
    blk.gdim = grid;
    blk.body = body;
    blk.dyn_smem = (char *)malloc(smem ? smem : 16);
    const int nwarps = (nthreads + 31) / 32;
    for (int t = 0; t < nthreads; t++) blk.fibers[t].stack = (char *)malloc(EMU_STACK);
    EmuBlock *saved = emu_blk;
    emu_blk = &blk;
    for (unsigned bx = 0; bx < grid.x; bx++) {
        blk.bid.x = bx;
        blk.bid.y = blk.bid.z = 0;
        blk.arrive.assign(1 + nwarps, 0);
        blk.gen.assign(1 + nwarps, 0);
        blk.xch.assign((size_t)nwarps * 32, 0);
        for (int t = 0; t < nthreads; t++) {
            EmuFiber &f = blk.fibers[t];
            f.tid.x = (unsigned)t;
            f.tid.y = f.tid.z = 0;
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = EMU_STACK;
            f.ctx.uc_link = &blk.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int live = nthreads;
        while (live > 0) {
            for (int t = 0; t < nthreads; t++) {
                if (blk.fibers[t].done) continue;
                blk.cur = t;
                swapcontext(&blk.sched, &blk.fibers[t].ctx);
                if (blk.fibers[t].done) live--;
            }
        }
    }
    emu_blk = saved;
    for (int t = 0; t < nthreads; t++) free(blk.fibers[t].stack);
    free(blk.dyn_smem);
}
