#!/usr/bin/env python3
"""Instrument a THROW-AWAY copy of the reference's figbird.cpp so that it dumps what it never prints.

TEST INFRASTRUCTURE ONLY (see oracle/figbird_oracle.h).  The reference has no tests, golden vectors
or probability dumps, so parity below the level of gapout.txt can only be pinned by
running the reference itself with taps added.  This script reads /root/reference/figbird.cpp,
inserts a few fprintf taps at anchor lines and writes the result to a temporary path given on the
command line; oracle/Makefile compiles that to oracle/_ref/figbird_dump.out and deletes the copy.
No reference source is ever written into the repository.

Taps (all written to the current directory of the instrumented run):
  fb_model.bin   the preamble's outputs = the hot path's constant inputs (format: FBM1, see below)
  fb_dump.txt    EST <gapfile> <estimate> <likelihood %a>      after the last EM iteration of each estimate
                 P <row> probs[5] counts[5] (%a)                rows of that estimate, if FB_DUMP_PROBS is set
                 BEST <gapfile> <estimate> <likelihood %a> <bestString>

fb_model.bin: "FBM1", int32 x6 (maxReadLength, maxInsertSize, thrMin, thrMax, gapProbCutOff,
maxDistance), then doubles: errorTypeProbs[25], errorPosDist[mrl], inPosDist[mrl], delPosDist[mrl],
insertLengthDistSmoothed[maxInsertSize].
"""
import sys

GLOBAL_ANCHOR = "FILE *gapOutFile;"
GLOBAL_TAP = "\nFILE *fbDump=NULL;\n"

EST_ANCHOR = "            if(likelihood>=maxLikelihood)\n"
EST_TAP = r'''
            if(fbDump)
            {
                fprintf(fbDump,"EST\t%s\t%d\t%a\n",gapFileName,gapEstimate,likelihood);
                if(getenv("FB_DUMP_PROBS"))
                {
                    for(long int fbi=0;fbi<endPos-startPos;fbi++)
                    {
                        fprintf(fbDump,"P\t%ld",fbi);
                        for(int fbj=0;fbj<5;fbj++)fprintf(fbDump,"\t%a",probsGap[fbi][fbj]);
                        for(int fbj=0;fbj<5;fbj++)fprintf(fbDump,"\t%a",countsGap[fbi][fbj]);
                        fprintf(fbDump,"\n");
                    }
                }
            }
'''

BEST_ANCHOR = "\t\tfinalize(maxGapEstimate);\n"
BEST_TAP = r'''
		if(fbDump)fprintf(fbDump,"BEST\t%s\t%d\t%a\t%s\n",gapFileName,maxGapEstimate,maxLikelihood,bestString);
'''

MODEL_ANCHOR = "    time_t now, end;\n"
MODEL_TAP = r'''
    {
        FILE *fbm=fopen("fb_model.bin","wb");
        int hdr[6]={maxReadLength,maxInsertSize,insertThresholdMin,insertThresholdMax,gapProbCutOff,maxDistance};
        fwrite("FBM1",1,4,fbm);
        fwrite(hdr,sizeof(int),6,fbm);
        fwrite(errorTypeProbs,sizeof(double),25,fbm);
        fwrite(errorPosDist,sizeof(double),maxReadLength,fbm);
        fwrite(inPosDist,sizeof(double),maxReadLength,fbm);
        fwrite(delPosDist,sizeof(double),maxReadLength,fbm);
        fwrite(insertLengthDistSmoothed,sizeof(double),maxInsertSize,fbm);
        fclose(fbm);
        fbDump=fopen("fb_dump.txt","w");
    }
'''

CLOSE_ANCHOR = "    fclose(gapInfoFile);\n    fclose(gapOutFile);\n    time(&end);"
CLOSE_TAP = "    if(fbDump)fclose(fbDump);\n"


def insert_before(src, anchor, tap):
    n = src.count(anchor)
    if n != 1:
        raise SystemExit(f"patch_ref: anchor occurs {n} times, expected 1: {anchor!r}")
    i = src.index(anchor)
    return src[:i] + tap + src[i:]


def insert_after(src, anchor, tap):
    n = src.count(anchor)
    if n != 1:
        raise SystemExit(f"patch_ref: anchor occurs {n} times, expected 1: {anchor!r}")
    i = src.index(anchor) + len(anchor)
    return src[:i] + tap + src[i:]


def main():
    if len(sys.argv) != 3:
        raise SystemExit("usage: patch_ref.py <reference figbird.cpp> <temporary output .cpp>")
    src = open(sys.argv[1], encoding="latin-1").read()
    src = insert_after(src, GLOBAL_ANCHOR, GLOBAL_TAP)
    src = insert_before(src, EST_ANCHOR, EST_TAP)
    src = insert_before(src, BEST_ANCHOR, BEST_TAP)
    src = insert_before(src, MODEL_ANCHOR, MODEL_TAP)
    src = insert_before(src, CLOSE_ANCHOR, CLOSE_TAP)
    open(sys.argv[2], "w", encoding="latin-1").write(src)


if __name__ == "__main__":
    main()
