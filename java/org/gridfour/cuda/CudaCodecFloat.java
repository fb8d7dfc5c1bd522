/* CudaCodecFloat.java -- drop-in for compress/CodecFloat.java behind the codec plugin API.  SOURCE ONLY.
 * The class lists both interfaces DIRECTLY because GvrsFileSpecification.addCompressionCodec inspects
 * getInterfaces() of the class itself (gvrs/GvrsFileSpecification.java:1608-1626); a public no-argument
 * constructor is invoked lazily by CodecHolder (gvrs/CodecHolder.java:189-234).
 */
package org.gridfour.cuda;

import org.gridfour.compress.ICompressionDecoder;
import org.gridfour.compress.ICompressionEncoder;

public class CudaCodecFloat extends CudaCodecBase implements ICompressionEncoder, ICompressionDecoder {

  public CudaCodecFloat() {
    super(G4Native.CODEC_FLOAT);
  }
}
