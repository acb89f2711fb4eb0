namespace {
int ls_plan(LsState &, const std::vector<uint32_t> &, const std::vector<uint32_t> &, const std::vector<uint32_t> &, bool, int)
{
    return fail(SEQA_ERR_UNSUPPORTED, "linear-space algorithms: not built yet");
}
int ls_run(seqa_ctx *, bool) { return fail(SEQA_ERR_UNSUPPORTED, "linear-space algorithms: not built yet"); }
void ls_release(LsState &) {}
} // namespace
