// cck_plan — command-line driver mirroring the reference's K-CBS CLI (demos/main.cpp:23-47) for the
// belief models, running the batched GPU low-level planner under the K-CBS search.
//   cck_plan -m maps/x.map -s scens/x.scen -k 2 [--solver K-CBS] [--lowlevel BSST] [--svc Blackmore]
//            [--pvc Blackmore] [-p 0.9] [-t 300] [--K 1024] [--seed 1] [--lltime 5] [--devices 1] [--prune] [--output plan.txt]
//            [--export DIR]   (the reference's exportBeliefPlan files: DIR/agentNN.txt, DIR/agentNN_covs.txt; DIR must exist)
// Prints one JSON line; --output writes the plan as text: per robot a line "robot r T W" followed by T rows.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "planner.hpp"

using namespace cck_b200;

int main(int argc, char** argv) {
    std::string map, scen, solver = "K-CBS", lowlevel = "BSST", svc = "Blackmore", pvc = "Blackmore", output, export_dir;
    int k = 1, K = 1024, devices = 1;
    double p_safe = 0.9, time_s = 300.0, lltime = 5.0;
    uint64_t seed = 1;
    bool prune = false;      // stock-SST pruning of old representatives (dead code in the reference's file, see PlannerParams::prune)
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "-m" || a == "--map") map = next();
        else if (a == "-s" || a == "--scen") scen = next();
        else if (a == "-k" || a == "--numAgents") k = std::atoi(next().c_str());
        else if (a == "--solver") solver = next();
        else if (a == "--lowlevel") lowlevel = next();
        else if (a == "--svc") svc = next();
        else if (a == "--pvc") pvc = next();
        else if (a == "-p" || a == "--p_safe") p_safe = std::atof(next().c_str());
        else if (a == "-t" || a == "--time") time_s = std::atof(next().c_str());
        else if (a == "--lltime") lltime = std::atof(next().c_str());
        else if (a == "--K") K = std::atoi(next().c_str());
        else if (a == "--seed") seed = std::strtoull(next().c_str(), nullptr, 10);
        else if (a == "--devices") devices = std::atoi(next().c_str());
        else if (a == "--prune") prune = true;
        else if (a == "-o" || a == "--output") output = next();
        else if (a == "--export") export_dir = next();
        else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (solver != "K-CBS" || lowlevel != "BSST") { std::fprintf(stderr, "only --solver K-CBS --lowlevel BSST are in scope\n"); return 2; }
    try {
        auto inst = std::make_shared<Instance>(map, scen, k, p_safe, svc, pvc);
        PlannerParams prm;
        prm.K = K; prm.seed = seed; prm.prune = prune;
        KCBS kcbs(inst, svc, pvc, prm, lltime, devices);
        KCBSResult r = kcbs.solve(time_s);
        int64_t launches = 0, cands = 0, nodes = 0;
        double ts = 0, tsel = 0, tex = 0, tcm = 0, txt = 0, tll = 0, cn = 0, cr = 0, cmi = 0; long long csurv = 0, cfresh = 0;
        for (int i = 0; i < k; ++i) {
            const PlannerStats& st = kcbs.planner(i).stats();
            launches += st.launches; cands += st.candidates; nodes += st.nodes;
            ts += st.t_sample; tsel += st.t_select; tex += st.t_expand; tcm += st.t_commit; txt += st.t_extract; tll += st.seconds;
            cn += st.c_witness; cr += st.c_replay; cmi += st.c_mirror; csurv += st.c_surv; cfresh += st.c_fresh;
        }
        // phase_seconds: summed over the robots' planners (which run concurrently, so they can exceed the wall clock)
        std::printf("{\"solved\": %s, \"seconds\": %.6f, \"cost\": %.6f, \"cbs_nodes\": %d, \"replans\": %d, \"launches\": %lld, \"candidates\": %lld, "
                    "\"tree_nodes\": %lld, \"obstacles\": %zu, \"agents\": %d, \"K\": %d, \"root_seconds\": %.6f, \"validate_seconds\": %.6f, "
                    "\"phase_seconds\": {\"low_level\": %.6f, \"sample\": %.6f, \"select\": %.6f, \"expand\": %.6f, \"commit\": %.6f, \"extract\": %.6f, "
                    "\"commit_witness_batch\": %.6f, \"commit_replay\": %.6f, \"commit_mirror\": %.6f, \"survivors\": %lld, \"new_witnesses\": %lld}}\n",
                    r.solved ? "true" : "false", r.seconds, r.cost, r.nodes_expanded, r.replans, (long long)launches, (long long)cands, (long long)nodes,
                    inst->getObstacles().size(), k, K, r.t_root, r.t_validate, tll, ts, tsel, tex, tcm, txt, cn, cr, cmi, csurv, cfresh);
        if (!export_dir.empty() && r.solved) exportBeliefPlan(r.paths, export_dir);
        if (!output.empty() && r.solved) {
            FILE* f = std::fopen(output.c_str(), "w");
            for (int i = 0; i < k; ++i) {
                const Trajectory& t = r.plan[i];
                std::fprintf(f, "robot %d %zu %zu\n", i, t.size(), t[0].v.size());
                for (const auto& s : t) { for (double v : s.v) std::fprintf(f, "%.17g ", v); std::fprintf(f, "\n"); }
            }
            std::fclose(f);
        }
        return r.solved ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "cck_plan: %s\n", e.what());
        return 3;
    }
}
