// bwt_convert_b200: format conversion (bwt_convert.cpp) with the host transcoders only. No GPU involved;
// exists so that the format code can be checked against the reference's bwt_convert on any machine.
#include <cstdlib>
#include <iostream>
#include <string>

#include <unistd.h>

#include "bwtm_host.hpp"

using namespace bwtm_host;

int main(int argc, char** argv)
{
  std::string input_format = "native", output_format = "native";
  int c = 0;
  while((c = getopt(argc, argv, "i:o:")) != -1)
  {
    switch(c)
    {
    case 'i': input_format = optarg; break;
    case 'o': output_format = optarg; break;
    default: std::exit(EXIT_FAILURE);
    }
  }
  if(argc - optind < 2 || !formatExists(input_format) || !formatExists(output_format))
  {
    std::cerr << "Usage: bwt_convert [-i format] [-o format] input output" << std::endl << std::endl;
    printFormats(std::cerr);
    std::exit(EXIT_FAILURE);
  }
  HostBWT bwt;
  if(!loadBWT(bwt, argv[optind], input_format)) { std::exit(EXIT_FAILURE); }
  std::cout << "BWT:              " << bwt.sequences << " sequences, " << bwt.bases << " bases, "
            << alphabetName(bwt.order()) << " alphabet" << std::endl;
  printSize("FMI", bwt.nativeSize(), bwt.bases);
  if(!serializeBWT(bwt, argv[optind + 1], output_format)) { std::exit(EXIT_FAILURE); }
  return 0;
}
