// fb_split — drop-in for the reference's SAM splitter (main.cpp / main.out): `fb_split <scaffolds.fa> [ignored]`
// reads result.sam in the working directory and writes gapInfo.txt, stat.txt, myout.sam, Gaps/gaps_<n>.sam,
// singletons.sam, linking.sam, unmapped.txt, scaffold.txt and counts.txt.  Same decisions as main.cpp
// (see csrc/host/fb_split.hpp); gap files are buffered and written at the end, so thousands of gaps do not need
// thousands of open descriptors (the reference keeps one FILE* per gap, main.cpp:1021-1034).
#include <cstdio>
#include <string>

#include "../csrc/host/fb_split.hpp"

int main(int argc, char** argv) {
    if (argc < 2) {
        printf("Invalid parameters\n");
        return 1;
    }
    fbhost::Scaffolds sc;
    if (!fbhost::read_scaffolds(argv[1], sc)) {
        printf("Can't open contig file\n");
        return 1;
    }
    fbhost::Splitter sp(sc);
    FILE* gi = fopen("gapInfo.txt", "w");
    sp.write_gap_info(gi);
    fclose(gi);
    FILE* probe = fopen("result.sam", "r");
    if (!probe) {
        printf("Can't open map file\n");
        return 1;
    }
    fclose(probe);
    FILE* myout = fopen("myout.sam", "w");
    FILE* linking = fopen("linking.sam", "w");
    FILE* single = fopen("singletons.sam", "w");
    FILE* un = fopen("unmapped.txt", "w");
    sp.run("result.sam", myout, single, linking, un, true);
    if (getenv("FB_TRACE")) fprintf(stderr, "[fb_trace] split: %ld stretches, %ld run again\n", sp.stretches, sp.reruns);
    fclose(myout);
    fclose(linking);
    fclose(single);
    fclose(un);
    for (size_t g = 0; g < sp.gaps.size(); g++) {
        std::string name = "Gaps/gaps_" + std::to_string(g) + ".sam";
        FILE* f = fopen(name.c_str(), "w");
        if (!f) {
            printf("Can't open %s\n", name.c_str());
            return 1;
        }
        fwrite(sp.gaps[g].text.data(), 1, sp.gaps[g].text.size(), f);
        fclose(f);
    }
    FILE* sf = fopen("scaffold.txt", "w");
    sp.write_scaffold_matrix(sf);
    fclose(sf);
    FILE* st = fopen("stat.txt", "w");
    sp.write_stat(st);
    fclose(st);
    FILE* cf = fopen("counts.txt", "w");
    sp.write_counts(cf);
    fclose(cf);
    return 0;
}
