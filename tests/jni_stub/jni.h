/* TEST-ONLY stand-in for <jni.h> (this image has no JDK): just enough of the JNI types and the function table for
 * `gcc -fsyntax-only java/jni/pb200_jni.c` to type-check the stub against include/pinot_b200.h (tests/test_abi.py). Never shipped. */
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef double jdouble; typedef int8_t jbyte; typedef int32_t jsize; typedef uint8_t jboolean;
typedef void* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jdoubleArray; typedef jarray jbyteArray; typedef jarray jobjectArray;
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_; typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*); void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*); void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
  jdouble* (*GetDoubleArrayElements)(JNIEnv*, jdoubleArray, jboolean*); void (*ReleaseDoubleArrayElements)(JNIEnv*, jdoubleArray, jdouble*, jint);
  jbyte* (*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*); void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint);
  jsize (*GetArrayLength)(JNIEnv*, jarray); jintArray (*NewIntArray)(JNIEnv*, jsize); jbyteArray (*NewByteArray)(JNIEnv*, jsize);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
  jclass (*FindClass)(JNIEnv*, const char*); jobjectArray (*NewObjectArray)(JNIEnv*, jsize, jclass, jobject);
  void (*SetObjectArrayElement)(JNIEnv*, jobjectArray, jsize, jobject); jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*); void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject); jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  jstring (*NewStringUTF)(JNIEnv*, const char*); jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  jboolean (*ExceptionCheck)(JNIEnv*); void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
};
